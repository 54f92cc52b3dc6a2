"""Which peer-copy path reaches NVLink speed?  (next-round probe for dissect_b200/exchange.py)

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 \
        tools/exp/p2p_copy_probe.py

Rank 0 writes 64 MB blocks into rank 1's buffer (mapped through torch's CUDA-IPC rebuild, as exchange.py does) on an
otherwise idle pair of GPUs, by
  a) torch `Tensor.copy_`                 (what torch does for cross-device copies: enables peer access itself)
  b) cudaMemcpyAsync, kind = DeviceToDevice (what exchange.py does today)
  c) cudaMemcpyPeerAsync with explicit device ordinals
  e) cudaMemcpyAsync into a torch symmetric-memory buffer (cuMem + peer access, what exchange.py maps now)
and prints GB/s per method plus every return code of cudaDeviceCanAccessPeer / cudaDeviceEnablePeerAccess.
Expected on an NVSwitch box: several hundred GB/s.  Measured in round 1 (2 x B200, NV18): a), b), c) all 2.6 ms per
64 MB = 25 GB/s, d) 0.025 ms = 2.6 TB/s; e) was added afterwards and has not run yet."""
import os
import sys
import time

import torch
import torch.distributed as dist


def main():
    from cuda.bindings import runtime as cudart
    from torch.multiprocessing.reductions import reduce_tensor
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("cpu:gloo,cuda:nccl", device_id=torch.device("cuda", local))
    nbytes = 64 << 20
    buf = torch.zeros(nbytes, dtype=torch.uint8, device="cuda")
    src = torch.full((nbytes,), rank + 1, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    handles = [None] * world
    dist.all_gather_object(handles, reduce_tensor(buf))
    peer_rank = (rank + 1) % world
    fn, args = handles[peer_rank]
    peer = fn(*args)
    pdev = peer.device.index
    print(f"[r{rank}] local device {local}, peer tensor on device {pdev}", file=sys.stderr)
    for a, b in ((local, pdev), (pdev, local)):
        err, can = cudart.cudaDeviceCanAccessPeer(a, b)
        print(f"[r{rank}] cudaDeviceCanAccessPeer({a},{b}) -> err {int(err)}, can {can}", file=sys.stderr)
    err = cudart.cudaDeviceEnablePeerAccess(pdev, 0)
    print(f"[r{rank}] cudaDeviceEnablePeerAccess({pdev}) -> {err}", file=sys.stderr)
    cudart.cudaGetLastError()
    stream = torch.cuda.Stream()
    d2d = cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice

    def timed(label, fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"[r{rank}] {label:44s} {dt * 1e3:8.3f} ms  {nbytes / dt / 1e9:8.1f} GB/s", file=sys.stderr)
        dist.barrier()

    def torch_copy():
        with torch.cuda.stream(stream):
            peer.copy_(src, non_blocking=True)

    def raw_async():
        (e,) = cudart.cudaMemcpyAsync(peer.data_ptr(), src.data_ptr(), nbytes, d2d, stream.cuda_stream)
        assert int(e) == 0, e

    def raw_peer():
        (e,) = cudart.cudaMemcpyPeerAsync(peer.data_ptr(), pdev, src.data_ptr(), local, nbytes, stream.cuda_stream)
        assert int(e) == 0, e

    def local_copy():
        (e,) = cudart.cudaMemcpyAsync(buf.data_ptr(), src.data_ptr(), nbytes, d2d, stream.cuda_stream)
        assert int(e) == 0, e

    timed("a) torch copy_ into the peer's IPC tensor", torch_copy)
    timed("b) cudaMemcpyAsync D2D (UVA) into the peer", raw_async)
    timed("c) cudaMemcpyPeerAsync into the peer", raw_peer)
    timed("d) cudaMemcpyAsync D2D, same device", local_copy)
    # e) the mapping dissect_b200/exchange.py uses now: torch symmetric memory (cuMem allocations + peer access)
    try:
        import torch.distributed._symmetric_memory as symm_mem
        st = symm_mem.empty(nbytes, dtype=torch.uint8, device=torch.device("cuda", local))
        st.zero_()
        torch.cuda.synchronize()
        hdl = symm_mem.rendezvous(st, dist.group.WORLD)
        ptrs = [int(x) for x in hdl.buffer_ptrs]
        print(f"[r{rank}] symmetric memory mapped: {len(ptrs)} buffers", file=sys.stderr)

        def symm_copy():
            (e,) = cudart.cudaMemcpyAsync(ptrs[peer_rank], src.data_ptr(), nbytes, d2d, stream.cuda_stream)
            assert int(e) == 0, e

        timed("e) cudaMemcpyAsync into the peer's symm_mem buffer", symm_copy)
        torch.cuda.synchronize()
        dist.barrier()
        print(f"[r{rank}] symm buffer holds {int(st[0])} / {int(st[-1])} (expected {(rank - 1) % world + 1})", file=sys.stderr)
    except Exception as ex:  # noqa: BLE001
        print(f"[r{rank}] symmetric memory path failed: {type(ex).__name__}: {ex}", file=sys.stderr)
    torch.cuda.synchronize()
    dist.barrier()
    # what arrived: the peer's last write must be visible here
    want = (rank - 1) % world + 1
    print(f"[r{rank}] buffer holds {int(buf[0])} / {int(buf[-1])} (expected {want})", file=sys.stderr)
    del peer
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
