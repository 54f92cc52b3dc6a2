"""Full-size correctness checks of a (multi-GPU) GRM build that need no n x n oracle: CHECKER code for bench.py,
tools/multi_gpu_check.py and the -m gpu tests (not product code; torch is used here for an independent fp64
restatement of the decode, genotype.cpp:735-738, 752-776, 922-959).

For a seeded vector w the engine's result must satisfy, row by row (Freivalds' matrix-product check in its vector
form, the size-independent analogue of the reference's test20 invariants, test.cpp:1722-1729):

    (N w)_i = sum_k q_ki (sum_j w_j q_kj)            exactly        w in {0,1}^n
    (S w)_i = sum_k z_ki (sum_j w_j z_kj)            to tolerance   w in {-1,+1}^n,   S = A o N as the caller gets them
    2 sum_{i>j} N_ij + sum_i N_ii = sum_k nn_k^2     exactly

The left sides are accumulated from the A and N TILES each rank received (its tile set), the right sides from the
packed .bed rows by plain torch matvecs; partial sums are all-reduced over the ranks.  A dropped, duplicated, stale or
misplaced 128 x 256 tile changes (S w)_i by about sqrt(256 / m) * m for 128 rows, 40x the tolerance at C2.
Per-SNP counts of the engine are compared with the torch restatement as well (exact).
"""
import math

import numpy as np
import torch


def _fields(pk_rows, n):
    """packed rows uint8 [r, >= ceil(n/4)] -> 2-bit patterns uint8 [r, n] (genotype.cpp:752, 784: LSB first)."""
    bpr = (n + 3) // 4
    shifts = torch.tensor([0, 2, 4, 6], dtype=torch.uint8, device=pk_rows.device)
    f = (pk_rows[:, :bpr, None] >> shifts[None, None, :]) & 3
    return f.reshape(pk_rows.shape[0], 4 * bpr)[:, :n]


def reference_matvecs(pk, keep, n, w_pm, w01, rank=0, world=1, rows_per_block=0):
    """Right-hand sides from this rank's share of the kept SNP rows (rows are dealt to ranks in contiguous runs).
    Returns dict(yS, yN, sum_nn2, freq (int64 [rows, 4] in the engine's order), row_index)."""
    dev = pk.device
    kept = torch.nonzero(torch.as_tensor(keep, device=dev) != 0).flatten() if keep is not None else torch.arange(pk.shape[0], device=dev)
    mk = kept.numel()
    lo, hi = mk * rank // world, mk * (rank + 1) // world
    mine = kept[lo:hi]
    rb = rows_per_block or max(32, int(4e8 // (8 * max(n, 1))))
    yS = torch.zeros(n, dtype=torch.float64, device=dev)
    yN = torch.zeros(n, dtype=torch.float64, device=dev)
    sum_nn2 = 0
    freqs = []
    code_val = torch.tensor([1., 0., 2., 3.], dtype=torch.float64, device=dev)   # pattern 00 / 01 / 10 / 11
    for b0 in range(0, mine.numel(), rb):
        idx = mine[b0: b0 + rb]
        f = _fields(pk[idx], n)
        c = torch.stack([(f == 1).sum(1), (f == 0).sum(1), (f == 2).sum(1), (f == 3).sum(1)], dim=1).to(torch.int64)
        freqs.append(c)
        nn = c[:, 1] + c[:, 2] + c[:, 3]
        sum_nn2 += int((nn * nn).sum().item())
        nnd = nn.to(torch.float64)
        p1 = (2. * c[:, 1].to(torch.float64) + c[:, 2].to(torch.float64)) / (2. * nnd)
        p2 = (2. * c[:, 3].to(torch.float64) + c[:, 2].to(torch.float64)) / (2. * nnd)
        sd = torch.sqrt(2. * p1 * (1. - p1))
        table = (code_val[None, :] - (1. + 2. * p2)[:, None]) / sd[:, None]
        table[:, 1] = 0.
        fl = f.to(torch.int64)
        Z = torch.gather(table, 1, fl)
        yS += (Z @ w_pm) @ Z
        del Z
        Q = (f != 1).to(torch.float64)
        yN += (Q @ w01) @ Q
        del Q, fl, f
    freq = torch.cat(freqs, 0) if freqs else torch.zeros((0, 4), dtype=torch.int64, device=dev)
    return {"yS": yS, "yN": yN, "sum_nn2": sum_nn2, "freq": freq, "rows": (lo, hi)}


def tile_matvecs(tile_rc, tiles_A, tiles_N, n, w_pm, w01, device, slab=512):
    """Left-hand sides from this rank's tiles.  tiles_*: numpy [ntiles, 256, 128] ([tile][col][row]), fp64, or a
    callable (t0, t1) -> that slab (so that a 100k-individual tile set never has to sit in host memory at once)."""
    n_pad = (n + 255) // 256 * 256
    wS = torch.zeros(n_pad, dtype=torch.float64, device=device)
    wN = torch.zeros(n_pad, dtype=torch.float64, device=device)
    wS[:n] = w_pm
    wN[:n] = w01
    yS = torch.zeros(n_pad, dtype=torch.float64, device=device)
    yN = torch.zeros(n_pad, dtype=torch.float64, device=device)
    sumN = torch.zeros((), dtype=torch.float64, device=device)
    rc = np.asarray(tile_rc).reshape(-1, 2)
    ar, ac = torch.arange(128, device=device), torch.arange(256, device=device)
    for t0 in range(0, rc.shape[0], slab):
        t1 = min(t0 + slab, rc.shape[0])
        r0 = torch.as_tensor(rc[t0:t1, 0].astype(np.int64) * 128, device=device)
        c0 = torch.as_tensor(rc[t0:t1, 1].astype(np.int64) * 256, device=device)
        A = torch.from_numpy(np.ascontiguousarray(tiles_A(t0, t1) if callable(tiles_A) else tiles_A[t0:t1])).to(device)
        N = torch.from_numpy(np.ascontiguousarray(tiles_N(t0, t1) if callable(tiles_N) else tiles_N[t0:t1])).to(device)
        gr = (r0[:, None] + ar[None, :])[:, None, :]          # [t, 1, 128]
        gc = (c0[:, None] + ac[None, :])[:, :, None]          # [t, 256, 1]
        valid = (gr < n) & (gc < n)
        lower = valid & (gr > gc)
        diag = valid & (gr == gc)
        N = torch.where(valid, N, torch.zeros_like(N))
        S = torch.where(valid, A, torch.zeros_like(A)) * N
        ridx = (r0[:, None] + ar[None, :])                    # [t, 128]
        cidx = (c0[:, None] + ac[None, :])                    # [t, 256]
        for M, w, y in ((S, wS, yS), (N, wN, yN)):
            L = torch.where(lower, M, torch.zeros_like(M))
            D = torch.where(diag, M, torch.zeros_like(M))
            wr, wc = w[ridx], w[cidx]
            y.index_add_(0, ridx.flatten(), (torch.einsum("tcr,tc->tr", L, wc) + torch.einsum("tcr,tc->tr", D, wc)).flatten())
            y.index_add_(0, cidx.flatten(), torch.einsum("tcr,tr->tc", L, wr).flatten())
        sumN += 2. * torch.where(lower, N, torch.zeros_like(N)).sum() + torch.where(diag, N, torch.zeros_like(N)).sum()
        del A, N, S
    return {"yS": yS[:n], "yN": yN[:n], "sumN": sumN}


def s_tolerance(n, m, operand="fp16"):
    """|(S w)_i - ref_i| / m allowed: 16 x the rms of sum_j w_j dA_ij for independent operand-rounding errors
    (rms dA = 3.5e-6 * sqrt(1e4 / m) for fp16 operands, 8x that for bf16).  Measured on B200 at
    C2 (n = 20000, m = 98107, fp16): max over the rows = 4.2 x that rms; a dropped 128 x 256 tile gives 320 x."""
    rms = {"fp16": 3.5e-6, "bf16": 2.8e-5, "fp16x2": 3.5e-6}[operand] * math.sqrt(1e4 / max(m, 1))
    return 16. * rms * math.sqrt(n)


def run(pk, keep, n, m_kept, tile_rc, tiles_A, tiles_N, engine_stats, rank, world, device, operand="fp16", seed=12345,
        dist=None, group=None):
    """All ranks call this; returns the result dict (identical on every rank)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    w01 = torch.randint(0, 2, (n,), generator=g).to(torch.float64).to(device)
    w_pm = (torch.randint(0, 2, (n,), generator=g).to(torch.float64) * 2. - 1.).to(device)
    ref = reference_matvecs(pk, keep, n, w_pm, w01, rank, world)
    lhs = tile_matvecs(tile_rc, tiles_A, tiles_N, n, w_pm, w01, device)
    # the engine's per-SNP counts (identical on every rank) against this rank's share of the torch restatement
    lo, hi = ref["rows"]
    stats_ok = bool(np.array_equal(np.asarray(engine_stats["freq"][lo:hi], dtype=np.int64), ref["freq"].cpu().numpy()))
    vec = torch.stack([ref["yS"], ref["yN"], lhs["yS"], lhs["yN"]])
    scal = torch.tensor([float(ref["sum_nn2"]), float(lhs["sumN"].item()), 1. if stats_ok else 0.], dtype=torch.float64,
                        device=device)
    if world > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(scal[:2], op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(scal[2:], op=dist.ReduceOp.MIN, group=group)
    yS_ref, yN_ref, yS, yN = vec[0], vec[1], vec[2], vec[3]
    n_bad = int((yN != yN_ref).sum().item())
    s_err = float(((yS - yS_ref).abs().max() / max(m_kept, 1)).item())
    tol = s_tolerance(n, m_kept, operand)
    out = {
        "N_freivalds_exact": n_bad == 0, "N_freivalds_rows_differing": n_bad,
        "N_sum_identity_exact": float(scal[0].item()) == float(scal[1].item()),
        "S_freivalds_max_err_over_m": s_err, "S_tol": tol, "S_ok": s_err <= tol and math.isfinite(s_err),
        "stats_counts_exact": bool(scal[2].item() == 1.),
        "what": f"(N w), (S w) row by row and sum(N) from the ranks' A and N tiles vs torch fp64 matvecs over the packed "
                f"rows; n={n}, m={m_kept}, seed {seed}",
    }
    out["ok"] = bool(out["N_freivalds_exact"] and out["N_sum_identity_exact"] and out["S_ok"] and out["stats_counts_exact"])
    return out
