"""Performance probe: device-resident synthetic .bed rows -> GRM, per-kernel-class device times."""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine  # noqa: E402
from dissect_b200.synth import synth_packed_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=8192)
    ap.add_argument("--m", type=int, default=32768)
    ap.add_argument("--cg", type=int, default=2)
    ap.add_argument("--operand", default="fp16")
    ap.add_argument("--count", default="mma")
    ap.add_argument("--chunk", type=int, default=8192)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--flat", action="store_true")
    a = ap.parse_args()
    pk, _ = synth_packed_cuda(a.n, a.m, 7, missing=0.01)
    torch.cuda.synchronize()
    eng = GrmEngine(operand=a.operand, count_method=a.count, chunk_snps=a.chunk, cta_group=a.cg, flat_norm=a.flat,
                    profile=True)
    eng.set_individuals(a.n)
    W = a.n * (a.n + 1) / 2 * a.m
    for s in range(a.steps):
        eng.begin(a.m)
        t0 = time.time()
        eng.push_bed_device(pk.data_ptr(), pk.stride(0), a.m)
        eng.finish()
        dt = time.time() - t0
        kt = eng.kernel_times()
        syrk_ms = kt["syrk"][0]
        print(json.dumps({"step": s, "cfg": vars(a), "wall_ms": dt * 1e3, "pair_snp_per_s": W / dt,
                          "syrk_tflops_alg": 2 * W / (syrk_ms * 1e-3) / 1e12 if syrk_ms else None,
                          "kernel_ms": {k: round(v[0], 3) for k, v in kt.items()},
                          "launches": {k: v[1] for k, v in kt.items()}}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
