"""LD mode throughput: SNP x SNP correlations of one SNP group, device-resident packed rows.
Prints one JSON line: pair*individual/s over the lower triangle (2 flop each in the S contraction), per-kernel times,
and the reference's own CPU time for a bounded sample (oracle/_ref/ldcorr) when it is built.

    python tools/gpu_ld_perf.py [--snps 16384] [--n 20000] [--steps 5]"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine, bedio  # noqa: E402
from dissect_b200.synth import synth_packed_cuda  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--snps", type=int, default=16384)
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--operand", default="fp16")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    pk, _ = synth_packed_cuda(a.n, a.snps, 9, maf_lo=0.05, missing=0.01)
    torch.cuda.synchronize()
    eng = GrmEngine(operand=a.operand, profile=True)
    W = a.snps * (a.snps + 1) / 2 * a.n

    def step():
        eng.ld_begin(a.n, a.snps)
        eng.push_bed_device(pk.data_ptr(), pk.stride(0), a.snps)
        eng.finish()

    for _ in range(3):
        step()
    eng.kernel_times()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    dt = e0.elapsed_time(e1) * 1e-3 / a.steps
    kt = eng.kernel_times()
    pairs, cnt = eng.ld_pairs(0.1, cap=16)
    out = {"metric": "ld_pair_individual_per_s", "value": W / dt, "unit": "pair*individual/s", "ms_per_step": 1e3 * dt,
           "config": {"workload": f"LD correlations of {a.snps} SNPs over {a.n} individuals, 1% missing (gwas.cpp:1156-1172)",
                      "operand": a.operand},
           "tflops_alg": 2 * W / dt / 1e12, "kernel_ms_per_step": {k: v[0] / a.steps for k, v in kt.items()},
           "step_includes": "geometry setup (dgrm_ld_begin allocates the m x m tile store), decode, expand, both contractions",
           "pairs_outside_0.1": cnt}
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ldcorr")
    if os.path.exists(exe) and not a.no_cpu:
        ns, ms = min(a.n, 4000), min(a.snps, 2000)     # bounded CPU sample
        with tempfile.TemporaryDirectory(prefix="dgrm_ld_") as wd:
            sub = bedio.synth_packed(ns, ms, 9, maf_lo=0.05, missing=0.01)
            bedio.write_bed(os.path.join(wd, "s"), sub, ns)
            t0 = time.time()
            r = subprocess.run([exe, "--bfile", os.path.join(wd, "s"), "--out", os.path.join(wd, "o")], capture_output=True, text=True,
                               env=dict(os.environ, LDCORR_THRESHOLD="0.1"))
            t = time.time() - t0
        if r.returncode == 0:
            out["cpu_baseline"] = {"value": ms * (ms + 1) / 2 * ns / t, "unit": "pair*individual/s", "kind": "reference",
                                   "cores": len(os.sched_getaffinity(0)),
                                   "sample": f"oracle/_ref/ldcorr (reference objects: load + normalizeGenotypes + two pdsyrk_ + "
                                             f"division + range scan) on {ms} SNPs x {ns} individuals, whole process {t:.2f} s"}
    print(json.dumps(out))
    eng.close()


if __name__ == "__main__":
    main()
