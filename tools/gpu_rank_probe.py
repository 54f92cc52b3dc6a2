"""One rank's share of a multi-GPU GRM on a single GPU (no exchange): per-rank kernel behaviour at large n.
    python tools/gpu_rank_probe.py --n 100000 --m 98304 --world 8 --rank 3 [--resident 3]
Run with DGRM_OVERLAP=0/1 to compare the two-stream decode/contract overlap."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine  # noqa: E402
from dissect_b200.synth import synth_packed_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100000)
    ap.add_argument("--m", type=int, default=98304)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=3)
    ap.add_argument("--resident", type=int, default=0)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--chunk", type=int, default=16384)
    a = ap.parse_args()
    pk, _ = synth_packed_cuda(a.n, a.m, 3, missing=0.01)
    eng = GrmEngine(rank=a.rank, world=a.world, chunk_snps=a.chunk, profile=True, resident_chunks=a.resident)
    eng.set_individuals(a.n)

    def step():
        eng.begin(a.m)
        eng.push_bed_device(pk.data_ptr(), pk.stride(0), a.m)
        eng.finish()

    step()
    eng.kernel_times()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    kt = eng.kernel_times()
    W = a.n * (a.n + 1) / 2 * a.m / a.world
    print(f"overlap={os.environ.get('DGRM_OVERLAP', '1')} resident={a.resident} n={a.n} m={a.m} rank {a.rank}/{a.world}: "
          f"{dt * 1e3:.1f} ms/step, {W / dt:.3e} pair*SNP/s per GPU, SYRK {2 * W / (kt['syrk'][0] / a.steps * 1e-3) / 1e12:.0f} TFLOP/s; "
          + ", ".join(f"{k} {v[0] / a.steps:.1f}" for k, v in kt.items() if v[0] > 0))
    eng.close()


if __name__ == "__main__":
    main()
