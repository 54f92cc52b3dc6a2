"""One isolated GPU parity case with verbose diagnostics (run by tools/gpu_ladder.py in a subprocess so a
trap in one configuration cannot poison the next).  Test infrastructure: compares libdgrm with the oracle."""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine, bedio  # noqa: E402
from oracle import grm_oracle as orc  # noqa: E402


def err_report(name, got, ref, mask):
    d = np.abs(got - ref)
    d[~mask] = 0
    bad = np.isnan(got) & mask
    d[bad] = np.inf
    mx = float(d.max())
    i, j = np.unravel_index(np.argmax(d), d.shape)
    print(f"  {name}: max|d|={mx:.3e} at ({i},{j}) got={got[i, j]!r} ref={ref[i, j]!r}  n_nan={int(bad.sum())}")
    return mx


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=300)
    ap.add_argument("--m", type=int, default=1000)
    ap.add_argument("--cg", type=int, default=1)
    ap.add_argument("--operand", default="fp16")
    ap.add_argument("--count", default="fp4")
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--missing", type=float, default=0.02)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--flat", action="store_true")
    ap.add_argument("--resident", type=int, default=0)
    ap.add_argument("--tolA", type=float, default=None)
    args = ap.parse_args()
    n, m = args.n, args.m
    pk = bedio.synth_packed(n, m, args.seed, missing=args.missing)
    codes = orc.decode_rows(pk, n)
    t0 = time.time()
    ref = orc.grm(codes, flat_norm=args.flat)
    t_ref = time.time() - t0
    eng = GrmEngine(operand=args.operand, count_method=args.count, chunk_snps=args.chunk, cta_group=args.cg,
                    flat_norm=args.flat, profile=True, resident_chunks=args.resident)
    eng.set_individuals(n)
    eng.begin(m)
    t0 = time.time()
    eng.push_bed_host(pk)
    eng.finish()
    t_gpu = time.time() - t0
    st = eng.snp_stats()
    ok = True
    for k in ("freq", "nn"):
        same = np.array_equal(st[k], ref["stats"][k])
        print(f"  stats {k}: {'bit-exact' if same else 'MISMATCH'}")
        ok &= same
    for k in ("p1", "p2", "sd"):
        same = np.array_equal(st[k].view(np.uint64), ref["stats"][k].view(np.uint64))
        print(f"  stats {k}: {'0 ulp' if same else 'MISMATCH max|d|=%g' % np.abs(st[k] - ref['stats'][k]).max()}")
        ok &= same
    low = np.tril(np.ones((n, n), bool))
    S = eng.fetch_block_cyclic("S", bs=(n, n))
    N = eng.fetch_block_cyclic("N", bs=(n, n))
    A = eng.fetch_block_cyclic("A", bs=(n, n))
    eN = err_report("N", N, ref["N"], low)
    eS = err_report("S", S, ref["S"], low)
    eA = err_report("A", A, ref["A"], low)
    dS = float(np.abs(np.diag(S) - np.diag(ref["S"])).max())
    print(f"  S diagonal max|d| = {dS:.3e}")
    diag_mean = float(np.mean(np.diag(ref["A"])))
    relA = eA / diag_mean
    tolA = args.tolA if args.tolA is not None else {"fp16": 2e-4, "bf16": 2e-3, "fp16x2": 4e-6}[args.operand]
    ok &= eN == 0.0 and relA <= tolA
    if eS > 1e-2 * m ** 0.5 * 10:
        # structure of the error: which 128x256 tiles / 8x8 blocks are off
        d = np.abs(np.where(low, S - ref["S"], 0))
        d[np.isnan(d)] = 1e30
        tiles = {}
        for r0 in range(0, n, 128):
            for c0 in range(0, n, 256):
                tiles[(r0 // 128, c0 // 256)] = float(d[r0:r0 + 128, c0:c0 + 256].max())
        print("  per-tile max err:", {k: f"{v:.2e}" for k, v in tiles.items()})
        sub = d[:16, :16]
        print("  err[0:16,0:16] > tol pattern:\n" + "\n".join("   " + "".join("X" if v > 1e-2 else "." for v in row) for row in sub))
    print(f"  kernel_times: {eng.kernel_times()}")
    print(json.dumps({"case": vars(args), "ok": bool(ok), "errN": eN, "errS": eS, "relA": relA, "t_gpu_s": t_gpu,
                      "t_oracle_s": t_ref}))
    eng.close()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
