"""CPU: the full-size checker itself (tools/fullsize_checks.py) against the numpy oracle — it must accept the
oracle's GRM cut into the engine's tile layout and reject a dropped, stale or misplaced tile and a wrong count."""
import numpy as np
import torch

from dissect_b200 import bedio
from dissect_b200.partition import tiles_of_rank
from oracle import grm_oracle as orc
from tools import fullsize_checks as fc


def _tiles_from_dense(M, n, coords):
    out = np.zeros((len(coords), 256, 128))
    for t, (rb, cp) in enumerate(coords):
        r0, c0 = rb * 128, cp * 256
        r1, c1 = min(r0 + 128, n), min(c0 + 256, n)
        if r0 < n and c0 < n:
            out[t, : c1 - c0, : r1 - r0] = M[r0:r1, c0:c1].T
    return out


def _case(n=700, m=900, seed=3):
    pk = bedio.synth_packed(n, m, seed, missing=0.03)
    ref = orc.grm(orc.decode_rows(pk, n))
    A = np.tril(ref["A"]) + np.tril(ref["A"], -1).T
    N = np.tril(ref["N"]) + np.tril(ref["N"], -1).T
    coords = tiles_of_rank(n, 0, 1)
    return pk, ref, coords, _tiles_from_dense(A, n, coords), _tiles_from_dense(N, n, coords)


def _run(pk, ref, coords, tA, tN, n, m, freq=None):
    stats = {"freq": ref["stats"]["freq"] if freq is None else freq}
    return fc.run(torch.from_numpy(pk), None, n, m, np.array(coords), tA, tN, stats, 0, 1, torch.device("cpu"))


def test_checker_accepts_the_oracle():
    n, m = 700, 900
    pk, ref, coords, tA, tN = _case(n, m)
    res = _run(pk, ref, coords, tA, tN, n, m)
    assert res["ok"], res
    assert res["S_freivalds_max_err_over_m"] < 1e-12          # exact inputs: only fp64 summation order differs


def test_checker_rejects_broken_tiles():
    n, m = 700, 900
    pk, ref, coords, tA, tN = _case(n, m)
    off = next(t for t, (rb, cp) in enumerate(coords) if rb * 128 >= cp * 256 + 256)    # a strictly-lower tile
    bad = tA.copy(); bad[off] = 0.                              # dropped tile
    assert not _run(pk, ref, coords, bad, tN, n, m)["S_ok"]
    other = next(t for t in range(len(coords)) if t != off and coords[t][0] * 128 >= coords[t][1] * 256 + 256)
    bad = tA.copy(); bad[off] = tA[other]                       # stale / misplaced tile
    assert not _run(pk, ref, coords, bad, tN, n, m)["S_ok"]
    badN = tN.copy(); badN[off, 5, 7] += 1                      # one pair count off by one
    r = _run(pk, ref, coords, tA, badN, n, m)
    assert not r["N_sum_identity_exact"] and not r["ok"]
    f = ref["stats"]["freq"].copy(); f[10, 2] += 1
    assert not _run(pk, ref, coords, tA, tN, n, m, freq=f)["stats_counts_exact"]


def test_checker_tolerates_operand_rounding_only():
    """fp16-sized noise on A passes, a 1e-3 relative error on one tile does not."""
    n, m = 700, 900
    pk, ref, coords, tA, tN = _case(n, m)
    rng = np.random.default_rng(1)
    noisy = tA + rng.normal(0, 3.5e-6 * np.sqrt(1e4 / m), tA.shape)
    assert _run(pk, ref, coords, noisy, tN, n, m)["S_ok"]
    off = next(t for t, (rb, cp) in enumerate(coords) if rb * 128 >= cp * 256 + 256)
    bad = tA.copy(); bad[off] += 5e-3
    assert not _run(pk, ref, coords, bad, tN, n, m)["S_ok"]
