"""Host-link probe for the e2e path (run under torchrun, one rank per GPU): plain pinned device->host and
host->device copies issued concurrently on the first k ranks, k = 1, 2, 4, ... world.  Answers whether the GPUs of
the box share one host link's worth of bandwidth: aggregate GB/s per k, per direction.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
        tools/d2h_concurrency_probe.py [--mb 1024] [--out profiles/r2_d2h_concurrency.json]
Also times the host-side widening the library does for streamed outputs (uint16 -> fp64 and fp32 -> fp64 with
non-temporal stores) as numpy-free C loops are not reachable from here: it reports numpy astype as a stand-in and says so."""
import argparse
import json
import os
import time

import numpy as np
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=4)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nbytes = args.mb << 20
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.fill_(3)
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    stream = torch.cuda.Stream()
    rows = []
    ks = sorted({k for k in (1, 2, 4, 8, world) if k <= world})
    for direction in ("d2h", "h2d", "both"):
        for k in ks:
            active = rank < k
            torch.cuda.synchronize()
            dist.barrier()
            t0 = time.perf_counter()
            if active:
                with torch.cuda.stream(stream):
                    for _ in range(args.reps):
                        if direction in ("d2h", "both"):
                            host.copy_(dev, non_blocking=True)
                        if direction == "h2d":
                            dev.copy_(host, non_blocking=True)
                if direction == "both":      # second stream: upload while downloading (full duplex?)
                    s2 = torch.cuda.Stream()
                    dev2 = getattr(main, "_dev2", None)
                    if dev2 is None:
                        main._dev2 = dev2 = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
                        main._host2 = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                    with torch.cuda.stream(s2):
                        for _ in range(args.reps):
                            dev2.copy_(main._host2, non_blocking=True)
                    s2.synchronize()
                stream.synchronize()
            dt = time.perf_counter() - t0
            t = torch.tensor([dt if active else 0.0], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            moved = k * args.reps * nbytes * (2 if direction == "both" else 1)
            rows.append({"direction": direction, "ranks_active": k, "seconds": float(t.item()),
                         "aggregate_GBps": moved / float(t.item()) / 1e9, "per_rank_GBps": moved / k / float(t.item()) / 1e9})
    # host widening stand-in: every rank converts 256 MB of uint16 to float64 at once
    src = np.ones(128 << 20, np.uint16)
    dst = np.empty(128 << 20, np.float64)
    for k in ks:
        dist.barrier()
        t0 = time.perf_counter()
        if rank < k:
            np.copyto(dst, src, casting="unsafe")
        dt = time.perf_counter() - t0
        t = torch.tensor([dt if rank < k else 0.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        rows.append({"direction": "host uint16->fp64 (numpy, 1 thread per rank)", "ranks_active": k, "seconds": float(t.item()),
                     "aggregate_GBps": k * (src.nbytes + dst.nbytes) / float(t.item()) / 1e9})
    if rank == 0:
        out = {"world": world, "mb_per_copy": args.mb, "reps": args.reps, "cpu_count": os.cpu_count(),
               "affinity": len(os.sched_getaffinity(0)), "rows": rows}
        print(json.dumps(out))
        for r in rows:
            print(f"{r['direction']:48s} ranks {r['ranks_active']}: aggregate {r['aggregate_GBps']:7.1f} GB/s")
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            json.dump(out, open(args.out, "w"), indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
