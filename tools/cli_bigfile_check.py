"""A GRM beyond the reference's 1 x 1-grid INT_MAX wall (n > 46 340) through the host CLI: synthesises a cohort,
writes .bed/.bim/.fam, runs dissect_b200/host/dgrm_make_grm (optionally with --gpu-engines G), and checks the
<out>.grm.dat it wrote: size, header, and a random 256-individual sub-block of A and N against the oracle
(statistics counted from the packed bytes of the whole cohort).  CHECKER / evidence tool, prints one JSON line.

    python tools/cli_bigfile_check.py --n 50000 --m 20000 --engines 1 --dir /tmp/c5cli"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import grm_oracle as orc  # noqa: E402
from tests.test_gpu_configs import stats_from_packed, sub_block_oracle  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=50000)
    ap.add_argument("--m", type=int, default=20000)
    ap.add_argument("--engines", type=int, default=1)
    ap.add_argument("--dir", default="/tmp/dgrm_cli_big")
    ap.add_argument("--keep-files", action="store_true")
    a = ap.parse_args()
    import torch
    from dissect_b200.synth import synth_packed_cuda
    n, m = a.n, a.m
    assert n % 4 == 0
    need = 14 + (n + 1) * n * 8 + m * n // 4
    os.makedirs(a.dir, exist_ok=True)
    free = shutil.disk_usage(a.dir).free
    if free < need * 1.1:
        print(json.dumps({"skipped": f"needs {need / 1e9:.1f} GB in {a.dir}, {free / 1e9:.1f} GB free"}))
        return
    pk, maf = synth_packed_cuda(n, m, 5, maf_lo=0.01, maf_hi=0.5, missing=0.01)
    host = pk.cpu().numpy()
    del pk
    torch.cuda.empty_cache()
    bf = os.path.join(a.dir, "in")
    with open(bf + ".bed", "wb") as f:
        f.write(bytes([0x6C, 0x1B, 0x01]))
        host.tofile(f)
    with open(bf + ".bim", "w") as f:
        f.write("".join(f"1\trs{k + 1}\t0\t{k + 1}\tA\tG\n" for k in range(m)))
    with open(bf + ".fam", "w") as f:
        f.write("".join(f"f{i + 1} i{i + 1} 0 0 0 -9\n" for i in range(n)))
    out = os.path.join(a.dir, "out")
    cli = os.path.join(ROOT, "dissect_b200", "host", "dgrm_make_grm")
    t0 = time.time()
    r = subprocess.run([cli, "--make-grm", "--bfile", bf, "--out", out, "--gpu-engines", str(a.engines)], capture_output=True, text=True)
    wall = time.time() - t0
    res = {"n": n, "m": m, "engines": a.engines, "cli_returncode": r.returncode, "cli_wall_s": wall}
    if r.returncode != 0:
        res["cli_tail"] = (r.stdout + r.stderr)[-1500:]
        print(json.dumps(res))
        sys.exit(1)
    dat = out + ".grm.dat"
    size = os.path.getsize(dat)
    res["grm_dat_bytes"] = size
    res["size_ok"] = size == 14 + (n + 1) * n * 8
    res["header_ok"] = open(dat, "rb").read(14) == orc.GRM_HEADER
    res["ids_ok"] = len(open(out + ".grm.ids").read().splitlines()) == n
    P = np.memmap(dat, dtype=np.float64, mode="r", offset=14, shape=(n, n + 1))      # P[c] = file column c
    sub = np.sort(np.random.default_rng(21).choice(n, 256, replace=False))
    st = stats_from_packed(host, n)
    S, Nn = sub_block_oracle(host, st, sub)
    A_ref = S * (1.0 / Nn)
    A = np.zeros((256, 256))
    N = np.zeros((256, 256))
    for j, c in enumerate(sub):
        col = np.asarray(P[c])
        A[j:, j] = col[sub[j:] + 1]           # P[r + 1, c] = A[r, c], r >= c
        N[j:, j] = [P[r][c] for r in sub[j:]]   # N[r, c] (r >= c) = N[c, r], stored in file column r at row c
    low = np.tril(np.ones((256, 256), bool))
    res["N_exact"] = bool(np.array_equal(N[low], Nn[low]))
    res["A_max_err"] = float(np.max(np.abs(A[low] - A_ref[low])))
    res["A_tol"] = 3e-4 * max(0.25, (1000.0 / m) ** 0.5)
    res["ok"] = bool(res["size_ok"] and res["header_ok"] and res["ids_ok"] and res["N_exact"] and res["A_max_err"] <= res["A_tol"])
    res["note"] = ("n x n = %.2e elements: beyond the INT_MAX local-matrix limit of a 1 x 1 grid (matrix.cpp:1773-1779); "
                   "kernel / N never existed as host matrices (options.lazyHostMatrices)" % (float(n) * n))
    del P
    if not a.keep_files:
        shutil.rmtree(a.dir, ignore_errors=True)
    print(json.dumps(res))
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
