"""GPU: BASELINE.json configs[3] and configs[4] geometries (50,000 individuals) against the oracle on a random
384-individual sub-block.  The oracle side needs the whole cohort's per-SNP statistics (counted here straight from the
packed bytes with a 256-entry table, compared bit for bit with the engine's) but only those individuals' columns.
SNP counts are reduced (the full-size runs are bench.py --config c4 / c5, which carry their own full-size checks);
everything that depends on the 50k geometry — 196 x 196 supertiles, 25 raster row groups, the resident cohort, the
saved (S, N) store and its views — is exercised at full width."""
import numpy as np
import pytest

from dissect_b200 import GrmEngine
from oracle import grm_oracle as orc
from tests.helpers import lower_mask

pytestmark = pytest.mark.gpu

N_IND = 50000
# chromosome lengths (Mb): configs[3]'s 22 partitions are proportional to them
CHR_MB = [248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64, 46, 50]


def count_table():
    t = np.zeros((4, 256), np.uint8)
    for b in range(256):
        for j in range(4):
            t[(b >> (2 * j)) & 3, b] += 1
    return t


def stats_from_packed(host, n):
    """per-SNP counts straight from the packed bytes (n % 4 == 0), then the reference's formulas (oracle.snp_stats)."""
    assert n % 4 == 0
    t = count_table()
    freq = np.zeros((host.shape[0], 4), np.int32)
    for r0 in range(0, host.shape[0], 2048):
        blk = host[r0:r0 + 2048]
        # pattern -> code: 00 -> 1 (hom A1), 01 -> 0 (missing), 10 -> 2, 11 -> 3
        for pat, code in ((0, 1), (1, 0), (2, 2), (3, 3)):
            freq[r0:r0 + blk.shape[0], code] = t[pat][blk].sum(axis=1, dtype=np.int64)
    nn = (freq[:, 1] + freq[:, 2] + freq[:, 3]).astype(np.int32)
    den = 2.0 * nn
    p1 = (2.0 * freq[:, 1] + freq[:, 2]) / den
    p2 = (2.0 * freq[:, 3] + freq[:, 2]) / den
    return {"freq": freq, "nn": nn, "p1": p1, "p2": p2, "sd": np.sqrt((2.0 * p1) * (1.0 - p1))}


def sub_codes(host, sub):
    pats = (host[:, sub // 4] >> (2 * (sub % 4)).astype(np.uint8)) & 3
    return orc.PATTERN_TO_CODE[pats]


def sub_block_oracle(host, st, sub, rows=slice(None)):
    codes = sub_codes(host[rows], sub)
    z = orc.standardise(codes, st["p2"][rows], st["sd"][rows])
    q = (codes != 0).astype(np.float64)
    return z.T @ z, q.T @ q


def check_stats(eng_st, st):
    assert np.array_equal(eng_st["freq"], st["freq"]) and np.array_equal(eng_st["nn"], st["nn"])
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(eng_st[k].view(np.uint64), st[k].view(np.uint64)), k


def synth(n, m, seed):
    import torch
    from dissect_b200.synth import synth_packed_cuda
    pk, maf = synth_packed_cuda(n, m, seed, maf_lo=0.001, maf_hi=0.5, missing=0.01)
    keep = (maf >= 0.01).cpu().numpy().astype(np.uint8)
    torch.cuda.synchronize()
    return pk, keep


def test_config5_geometry_subset_vs_oracle():
    """configs[4]: 50,000 individuals, 1 % missing, MAF-filter extract list (SNPs cut from 200,000 to 24,000)."""
    n, m = N_IND, 24000
    pk, keep = synth(n, m, 5)
    eng = GrmEngine()
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_bed_device(pk.data_ptr(), pk.stride(0), m, row_keep=keep)
    eng.finish()
    host = pk.cpu().numpy()[keep != 0]
    del pk
    st = stats_from_packed(host, n)
    check_stats(eng.snp_stats(), st)
    sub = np.sort(np.random.default_rng(11).choice(n, 384, replace=False))
    S, Nn = sub_block_oracle(host, st, sub)
    eng.set_output_subset(sub)
    low = lower_mask(384)
    N = eng.fetch_block_cyclic("N", bs=(384, 384))
    A = eng.fetch_block_cyclic("A", bs=(384, 384))
    assert np.array_equal(N[low], Nn[low])
    err = float(np.max(np.abs(A[low] - (S * (1.0 / Nn))[low])))
    assert err <= 3e-5, err                       # fp16 operands at m = 2.4e4 (1e-5 is met from m = 1e5, configs[1] test)
    eng.close()


def test_config4_geometry_groups_and_loco_views_vs_oracle():
    """configs[3]: 50,000 individuals, 22 chromosome partitions from ONE resident cohort (SNPs cut from 300,000 to
    26,400): the global GRM is kept as (S, N), three of the partitions are built from the same resident rows, and both
    the partition GRM and its leave-one-chromosome-out view (global - partition, singlereml.cpp:298) are checked."""
    n, m = N_IND, 26400
    pk, keep = synth(n, m, 4)
    edges = np.concatenate([[0], np.round(np.cumsum(CHR_MB) / np.sum(CHR_MB) * m).astype(np.int64)])
    eng = GrmEngine()
    eng.set_individuals(n)
    eng.store_bed_device(pk.data_ptr(), pk.stride(0), m)
    host_all = pk.cpu().numpy()
    del pk
    kept = np.nonzero(keep)[0]
    host = host_all[kept]
    st = stats_from_packed(host, n)
    sub = np.sort(np.random.default_rng(12).choice(n, 384, replace=False))
    low = lower_mask(384)
    eng.begin(len(kept))
    eng.push_stored(0, m, row_keep=keep)
    eng.finish()
    check_stats(eng.snp_stats(), st)
    eng.save("copy")
    S_all, N_all = sub_block_oracle(host, st, sub)
    for g in (0, 9, 21):
        g0, g1 = int(edges[g]), int(edges[g + 1])
        rows = slice(int(np.searchsorted(kept, g0)), int(np.searchsorted(kept, g1)))
        eng.begin(max(1, int(keep[g0:g1].sum())))
        eng.push_stored(g0, g1 - g0, row_keep=keep[g0:g1])
        eng.finish()
        S_g, N_g = sub_block_oracle(host, st, sub, rows)
        eng.set_output_subset(sub)
        mg = rows.stop - rows.start
        N = eng.fetch_block_cyclic("N", bs=(384, 384))
        A = eng.fetch_block_cyclic("A", bs=(384, 384))
        assert np.array_equal(N[low], N_g[low]), g
        assert np.max(np.abs(A[low] - (S_g * (1.0 / N_g))[low])) <= 3e-4 * max(0.25, (1000.0 / mg) ** 0.5), g
        eng.set_view("saved-current")                 # leave this chromosome out
        N = eng.fetch_block_cyclic("N", bs=(384, 384))
        A = eng.fetch_block_cyclic("A", bs=(384, 384))
        assert np.array_equal(N[low], (N_all - N_g)[low]), g
        assert np.max(np.abs(A[low] - ((S_all - S_g) * (1.0 / (N_all - N_g)))[low])) <= 3e-5, g
        eng.set_view("current")
    eng.close()


def test_config3_geometry_subset_vs_oracle_whole_and_one_of_eight_engines():
    """configs[2]: 100,000 individuals (391 x 391 supertiles, 49 raster row groups; a 40 GB fp64 + 20 GB int32 tile store
    on one GPU), SNPs cut from 500,000 to 8,192.  Checked on a random 384-individual sub-block for (a) the whole
    triangle on one engine and (b) engine 5 of an 8-engine partition — the compact macro-block raster at its real size,
    with operands expanded only for the individuals that engine's tiles touch — on the entries it owns."""
    n, m = 100000, 8192
    pk, keep = synth(n, m, 3)
    host = pk.cpu().numpy()[keep != 0]
    st = stats_from_packed(host, n)
    sub = np.sort(np.random.default_rng(13).choice(n, 384, replace=False))
    S, Nn = sub_block_oracle(host, st, sub)
    A_ref = S * (1.0 / Nn)
    low = lower_mask(384)
    tol = 3e-4 * max(0.25, (1000.0 / host.shape[0]) ** 0.5)
    for rank, world in ((0, 1), (5, 8)):
        eng = GrmEngine(rank=rank, world=world)
        eng.set_individuals(n)
        eng.begin(m)
        eng.push_bed_device(pk.data_ptr(), pk.stride(0), m, row_keep=keep)
        eng.finish()
        if world == 1:
            check_stats(eng.snp_stats(), st)
        owned = np.zeros((384, 384), bool)
        have = {(int(rb), int(cp)) for rb, cp in eng.tile_coords()}
        for i, r in enumerate(sub):
            for j, c in enumerate(sub[: i + 1]):
                owned[i, j] = (r // 128, c // 256) in have
        assert owned.any()
        if world == 1:
            assert owned[low].all()
        eng.set_output_subset(sub)
        N = eng.fetch_block_cyclic("N", bs=(384, 384), fill=-5.0)
        A = eng.fetch_block_cyclic("A", bs=(384, 384), fill=-5.0)
        assert np.array_equal(N[owned], Nn[owned]), (rank, world)
        assert np.max(np.abs(A[owned] - A_ref[owned])) <= tol, (rank, world)
        assert np.all(N[low & ~owned] == -5.0)            # entries of other engines' tiles are left alone
        eng.close()
    del pk
