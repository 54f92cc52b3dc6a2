"""Packed-chunk exchange between the GPUs of one box without SM work and without host synchronisation:
a peer-to-peer copy-engine all-gather ordered by stream memory operations.

The multi-GPU path has one exchange step per SNP chunk: every rank holds 1/G of the chunk's packed
2-bit rows and needs all of them.  `torch.distributed.all_gather_into_tensor` (NCCL) does that with a kernel — and a
kernel cannot be co-scheduled with the persistent tcgen05 contraction, which holds every SM (168 registers x 256
threads and 132 KB of shared memory per CTA; NCCL's all-gather CTAs need 96 registers x up to 640 threads), so the
all-gather of chunk c+1 only ran in the gaps between contraction launches (measured with `DGRM_TRACE=1`: the next
chunk's rows arrived exactly when the running SYRK launch retired).  Copy engines need no SM:

  * every rank maps every peer's gather buffers and flag words (CUDA IPC) and writes its shard of chunk s straight
    into slot `rank` of buffer s % 3 on every rank (`cudaMemcpyAsync`, device to device over NVLink), followed on
    the same stream by a 4-byte copy of s into the peer's `arrived[rank]` word;
  * the consumer's stream waits with `cuStreamWaitValue32(arrived[q] >= s)` for every q, then the engine's row
    gather reads the buffer (`dgrm_push_bed_device` with that stream as producer);
  * when the gather has consumed buffer s % 3 (`dgrm_release_inputs`), the rank writes s into every peer's
    `consumed[rank]` word, and a producer waits for `consumed[p] >= s - 3` before it overwrites p's buffer.

Everything is stream-ordered: the host never blocks and there is no barrier; the sequence numbers are global and
monotonic, so buffers are reused across GRMs without draining.

STATUS (round 2): the default exchange of dissect_b200/multi.py (`exchange="auto"`).  The peers' buffers are mapped
through torch symmetric memory (cuMem allocations with peer access, the NVLink path NCCL itself uses); parity-green at
2 GPUs (tools/multi_gpu_check.py --exchange p2p: tiles vs oracle, block-cyclic egress, parallel writer) and faster than
NCCL: 38.15 vs 40.10 ms per step at 2 GPUs (profiles/r2_bench_n2_p2p_symm.json vs r2_bench_n2_xg1.json), i.e. the
2-GPU step costs what the 1-GPU step costs.  Legacy CUDA-IPC mappings (`DGRM_EXCHANGE_MAP=ipc`) also pass parity but
copy at only ~25 GB/s on this platform (52.7 ms per step): kept for boxes without symmetric memory support.
"""
import os

import torch
import torch.distributed as dist

NBUF = 3
_SEQ_CAP = 1 << 20


def _ck(res):
    err = res[0] if isinstance(res, tuple) else res
    if int(err) != 0:
        raise RuntimeError(f"CUDA call failed: {err}")
    return res[1:] if isinstance(res, tuple) and len(res) > 1 else None


class PeerExchange:
    def __init__(self, rows_per_rank, pitch, rank, world, host_group):
        """rows_per_rank: shard rows every rank contributes per chunk; pitch: bytes per row (16-byte multiple).
        host_group: a gloo process group, used once to exchange the IPC handles."""
        from cuda.bindings import driver as cu
        from cuda.bindings import runtime as cudart
        self.cu, self.cudart = cu, cudart
        self.rank, self.world, self.shard, self.pitch = rank, world, rows_per_rank, pitch
        self.host_group = host_group
        self.error = None
        self._peer_tensors = []      # keep the mappings alive
        self.peer_buf, self.peer_flag = [], []
        self.mapping = os.environ.get("DGRM_EXCHANGE_MAP", "symm")
        self.seqsrc = torch.arange(_SEQ_CAP, dtype=torch.int32, device="cuda")
        nbuf_bytes = (world * rows_per_rank * pitch + 4095) // 4096 * 4096
        try:
            if self.mapping == "symm":
                # torch symmetric memory: cuMem (VMM) allocations exchanged as file descriptors and mapped with peer
                # access -- the path NCCL itself uses, direct over NVLink.  (Legacy cudaIpc mappings, below, move only
                # ~25 GB/s on this platform even between idle GPUs: tools/exp/p2p_copy_probe.py.)
                import torch.distributed._symmetric_memory as symm_mem
                dev = torch.device("cuda", torch.cuda.current_device())
                self._symm = symm_mem.empty(NBUF * nbuf_bytes + 4096, dtype=torch.uint8, device=dev)
                self._symm.zero_()
                torch.cuda.synchronize()
                self._hdl = symm_mem.rendezvous(self._symm, dist.group.WORLD)
                base = [int(x) for x in self._hdl.buffer_ptrs]
                nrow = world * rows_per_rank
                self.bufs = [self._symm[k * nbuf_bytes: k * nbuf_bytes + nrow * pitch].view(nrow, pitch) for k in range(NBUF)]
                self.flags = self._symm[NBUF * nbuf_bytes: NBUF * nbuf_bytes + 8 * world].view(torch.int32)
                for p in range(world):
                    self.peer_buf.append([base[p] + k * nbuf_bytes for k in range(NBUF)])
                    self.peer_flag.append(base[p] + NBUF * nbuf_bytes)
            else:
                from torch.multiprocessing.reductions import reduce_tensor
                self.bufs = [torch.zeros((world * rows_per_rank, pitch), dtype=torch.uint8, device="cuda") for _ in range(NBUF)]
                self.flags = torch.zeros(2 * world, dtype=torch.int32, device="cuda")   # arrived[world] | consumed[world]
                torch.cuda.synchronize()
                mine = [reduce_tensor(t) for t in self.bufs + [self.flags]]
                everyone = [None] * world
                dist.all_gather_object(everyone, mine, group=host_group)
                for p in range(world):
                    ts = self.bufs + [self.flags] if p == rank else [fn(*args) for (fn, args) in everyone[p]]
                    self._peer_tensors.append(ts)
                    self.peer_buf.append([t.data_ptr() for t in ts[:NBUF]])
                    self.peer_flag.append(ts[NBUF].data_ptr())
                    if p != rank:
                        dev = ts[0].device.index
                        if not torch.cuda.can_device_access_peer(torch.cuda.current_device(), dev):
                            raise RuntimeError(f"no peer access from device {torch.cuda.current_device()} to {dev}")
                        cudart.cudaDeviceEnablePeerAccess(dev, 0)    # "already enabled" is fine
                        cudart.cudaGetLastError()
        except Exception as ex:  # noqa: BLE001 - every rank must still reach the agreement below
            self.error = f"{type(ex).__name__}: {ex}"
        # all ranks agree on whether the mapping worked everywhere (callers fall back to NCCL otherwise)
        okt = torch.tensor([0 if self.error else 1], dtype=torch.int32)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN, group=host_group)
        self.ok = bool(okt.item())
        self.copy_stream = torch.cuda.Stream()
        self.arrive_stream = torch.cuda.Stream()
        self.rel_stream = torch.cuda.Stream()
        self.seq_post = 0
        self.seq_cons = 0
        self._trace = [] if os.environ.get("DGRM_TRACE") else None
        self._geq = cu.CUstreamWaitValue_flags.CU_STREAM_WAIT_VALUE_GEQ
        self._d2d = cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice
        dist.barrier(group=host_group)

    # -- helpers ------------------------------------------------------------------------------------------
    def _wait_geq(self, stream, addr, value):
        _ck(self.cu.cuStreamWaitValue32(stream.cuda_stream, addr, value, self._geq))

    def _copy(self, stream, dst, src, nbytes):
        _ck(self.cudart.cudaMemcpyAsync(dst, src, nbytes, self._d2d, stream.cuda_stream))

    # -- protocol -----------------------------------------------------------------------------------------
    def post(self, src_ptr, nrows):
        """Send my shard of the next chunk (contiguous rows of `pitch` bytes at device address src_ptr, readable in
        copy_stream order) to every rank.  Returns the chunk's sequence number."""
        self.seq_post += 1
        s = self.seq_post
        assert s < _SEQ_CAP
        trace = self._trace is not None and len(self._trace) < 64
        if trace:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.copy_stream)
        k = s % NBUF
        off = self.rank * self.shard * self.pitch
        for d in range(self.world):
            p = (self.rank + d) % self.world              # stagger the targets over the NVLink ports
            if s > NBUF:                                  # p's engine has read the chunk that used this buffer
                self._wait_geq(self.copy_stream, self.flags.data_ptr() + 4 * (self.world + p), s - NBUF)
            if nrows > 0:
                self._copy(self.copy_stream, self.peer_buf[p][k] + off, src_ptr, nrows * self.pitch)
            self._copy(self.copy_stream, self.peer_flag[p] + 4 * self.rank, self.seqsrc.data_ptr() + 4 * s, 4)
        if trace:
            e1.record(self.copy_stream)
            self._trace.append((s, nrows, e0, e1))
        return s

    def consume(self, engine, row_keep):
        """Hand the next chunk to the engine once every rank's shard has landed (stream-ordered, no host wait)."""
        self.seq_cons += 1
        s = self.seq_cons
        k = s % NBUF
        for q in range(self.world):
            self._wait_geq(self.arrive_stream, self.flags.data_ptr() + 4 * q, s)
        engine.push_bed_device(self.bufs[k].data_ptr(), self.pitch, self.world * self.shard, row_keep=row_keep,
                               stream=self.arrive_stream.cuda_stream)
        engine.release_inputs(self.rel_stream.cuda_stream)
        for d in range(self.world):
            p = (self.rank + d) % self.world
            self._copy(self.rel_stream, self.peer_flag[p] + 4 * (self.world + self.rank), self.seqsrc.data_ptr() + 4 * s, 4)
        return s

    def close(self):
        torch.cuda.synchronize()
        if self._trace:
            import sys
            for (s, nrows, e0, e1) in self._trace[:: max(1, len(self._trace) // 16)]:
                print(f"[exchange r{self.rank}] chunk {s}: {nrows} rows x {self.world} targets, "
                      f"{e0.elapsed_time(e1):.3f} ms on the copy stream (waits included)", file=sys.stderr)
        dist.barrier(group=self.host_group)     # nobody unmaps while a peer may still be writing
        self._peer_tensors = None
        self.peer_buf = self.peer_flag = None
        self._hdl = None
        dist.barrier(group=self.host_group)
