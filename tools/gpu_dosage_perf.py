"""Dosage ingest throughput: real-valued genotype rows (what the reference's BGEN reader forms
per variant) through dgrm_push_dosage_host to a finished GRM.  Prints one JSON line: genotype*individual pair updates
per second end to end (host rows -> finished S, N on the device), the per-kernel times, and the HBM roofline fraction
of the dosage kernels (algorithmic bytes per genotype: mask fold 1 byte read; statistics two passes of 8 bytes --
timed together as "count"; expand 8 bytes read, 2 (operand) + 0.5 (fp4 count operand) bytes written).

    python tools/gpu_dosage_perf.py [--n 20000] [--snps 16384] [--steps 3]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=20000)
    ap.add_argument("--snps", type=int, default=16384)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--missing", type=float, default=0.01)
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    f = rng.uniform(0.05, 0.5, size=(a.snps, 1)).astype(np.float32)
    d = (rng.random((a.snps, a.n), dtype=np.float32) < f).astype(np.float64)
    d += (rng.random((a.snps, a.n), dtype=np.float32) < f)
    d = np.clip(d + rng.integers(-25, 26, size=(a.snps, a.n), dtype=np.int8) / 255.0, 0.0, 2.0)
    miss = (rng.random((a.snps, a.n), dtype=np.float32) < a.missing).astype(np.uint8)
    eng = GrmEngine(profile=True)
    eng.set_individuals(a.n)

    def step():
        eng.begin(a.snps)
        eng.push_dosage_host(d, miss)
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
    g = float(a.snps) * a.n
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs_burst", peaks.get("hbm_gbs", 0)) or 0) or 6550.0
    ms = {k: v[0] / a.steps for k, v in kt.items()}
    stats_bytes, expand_bytes = g * (2 * 8 + 1), g * (8 + 2.5)
    out = {"metric": "dosage_ingest", "n": a.n, "snps": a.snps, "ms_per_step_wall": 1e3 * dt,
           "pair_updates_per_s": a.n * (a.n + 1) / 2 * a.snps / dt,
           "h2d_bytes_per_step": int(g * 9), "kernel_ms_per_step": ms,
           "roofline": {"bound": "hbm", "peak": peak, "unit": "GB/s",
                        "stats": {"achieved": stats_bytes / (ms["count"] * 1e-3) / 1e9 if ms["count"] else None},
                        "expand": {"achieved": expand_bytes / (ms["expand"] * 1e-3) / 1e9 if ms["expand"] else None}},
           "note": "wall time includes the host copy of the caller's rows into pinned staging and the H2D of 9 bytes per genotype"}
    for k in ("stats", "expand"):
        v = out["roofline"][k]["achieved"]
        out["roofline"][k]["frac"] = v / peak if v else None
    print(json.dumps(out))

    # the same genotypes as inflated BGEN layout-2 blocks with 8-bit probabilities, unpacked on the device
    del d
    blk = 10 + a.n + 2 * a.n
    data = np.zeros((a.snps, blk), dtype=np.uint8)
    data[:, 0:4] = np.frombuffer(np.uint32(a.n).tobytes(), dtype=np.uint8)
    data[:, 4], data[:, 6], data[:, 7] = 2, 2, 2
    data[:, 8:8 + a.n] = 2 | (miss << 7)
    data[:, 9 + a.n] = 8
    data[:, 10 + a.n:] = rng.integers(0, 128, size=(a.snps, 2 * a.n), dtype=np.uint8)
    offsets = np.arange(a.snps + 1, dtype=np.int64) * blk
    flat = data.reshape(-1)

    def step_raw():
        eng.begin(a.snps)
        eng.push_bgen_blocks_host(flat, offsets, 2, a.n)
        eng.finish()

    step_raw()
    eng.kernel_times()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        step_raw()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.steps
    ms = {k: v[0] / a.steps for k, v in eng.kernel_times().items()}
    print(json.dumps({"metric": "bgen_block_ingest", "n": a.n, "snps": a.snps, "bits": 8, "ms_per_step_wall": 1e3 * dt,
                      "pair_updates_per_s": a.n * (a.n + 1) / 2 * a.snps / dt, "h2d_bytes_per_step": int(flat.nbytes),
                      "kernel_ms_per_step": ms,
                      "unpack": {"bytes_per_genotype": 3 + 8, "achieved_GBps": g * 11 / (ms["gather"] * 1e-3) / 1e9 if ms["gather"] else None},
                      "note": "gather = bgen_unpack_kernel (3 bytes read, 8 written per genotype); count / expand as above"}))


if __name__ == "__main__":
    main()
