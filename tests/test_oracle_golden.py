"""CPU: the oracle restatements (numpy + plain C) against what the REFERENCE ITSELF produced
(tests/golden, written by oracle/make_golden.py running oracle/_ref built from the unmodified sources)."""
import os

import numpy as np
import pytest

from oracle import c_oracle
from oracle import grm_oracle as orc
from tests.helpers import GOLDEN_GRM_CASES, kept_rows, load_case, lower_mask


@pytest.mark.parametrize("name", GOLDEN_GRM_CASES)
def test_numpy_oracle_matches_reference(name):
    c = load_case(name)
    codes = orc.decode_rows(kept_rows(c), c["n_fam"], c["keep_mask"])
    g = orc.grm(codes, flat_norm=c["flat"])
    st, ref = g["stats"], c["ref_stats"]
    # integer counts bit-exact, statistics 0 ulp
    assert np.array_equal(st["freq"], ref["freq"])
    assert np.array_equal(st["nn"], ref["nn"])
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(st[k].view(np.uint64), ref[k].view(np.uint64)), k
    n = c["n"]
    low = lower_mask(n)
    assert np.array_equal(g["N"][low], c["ref_N"][low])                       # exact integers
    # A: only the dsyrk summation order differs (OpenBLAS vs numpy) -> ~1e-15 absolute
    assert np.max(np.abs(g["A"][low] - c["ref_A"][low])) <= 1e-13
    # the packed file layout (kernel.cpp:944-974, matrix.cpp:2324-2334): same size/header, N block identical
    blob = orc.pack_grm_dat(g["A"], g["N"])
    assert len(blob) == len(c["ref_dat"]) and blob[:14] == c["ref_dat"][:14]
    mine = np.frombuffer(blob, np.float64, offset=14)
    theirs = np.frombuffer(c["ref_dat"], np.float64, offset=14)
    assert np.max(np.abs(mine - theirs)) <= 1e-13


def test_standardised_genotypes_bit_exact():
    c = load_case("tiny7_m40")
    codes = orc.decode_rows(kept_rows(c), c["n_fam"], c["keep_mask"])
    g = orc.grm(codes)
    z = np.fromfile(os.path.join(c["dir"], "ref.z.bin")).reshape(2, codes.shape[0], codes.shape[1])
    assert np.array_equal(z[0].view(np.uint64), g["Z"].view(np.uint64))      # normalizeGenotypes, 0 ulp
    assert np.array_equal(z[1], (codes != 0).astype(np.float64))              # missings matrix


@pytest.mark.parametrize("name", GOLDEN_GRM_CASES)
def test_c_restatement_matches_reference(name):
    if not c_oracle.available():
        pytest.skip("oracle/_ref/libgrm_oracle.so not built (python -c 'import __graft_entry__ as g; g.build()')")
    c = load_case(name)
    g = c_oracle.grm(kept_rows(c), c["n_fam"], c["keep_mask"], flat_norm=c["flat"])
    ref = c["ref_stats"]
    assert np.array_equal(g["stats"]["freq"], ref["freq"]) and np.array_equal(g["stats"]["nn"], ref["nn"])
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(g["stats"][k].view(np.uint64), ref[k].view(np.uint64)), k
    low = lower_mask(c["n"])
    assert np.array_equal(g["N"][low], c["ref_N"][low])
    assert np.max(np.abs(g["A"][low] - c["ref_A"][low])) <= 1e-13


def test_badsnp_fixture():
    c = load_case("badsnp")
    codes = orc.decode_rows(c["packed"], c["n_fam"])
    st = orc.snp_stats(codes)
    with pytest.raises(ValueError) as ei:
        orc.standardise(codes, st["p2"], st["sd"])
    bad = [c["bim"][i] for i in ei.value.args[0][1]]
    assert bad == open(os.path.join(c["dir"], "ref.badsnps")).read().split()   # what the reference wrote


def test_sanitize_scan_order_and_reject():
    c = load_case("sanitize_reject")
    codes = orc.decode_rows(c["packed"], c["n_fam"])
    g = orc.grm(codes)
    pairs = orc.low_overlap_pairs(g["N"], 10.0)
    assert len(pairs) == 7 * 6 // 2          # every pair overlaps on < 10 SNPs -> the reference skipped the GRM
    assert c["meta"]["grm_written"] == "0"


def test_block_cyclic_helpers():
    G = np.arange(49, dtype=np.float64).reshape(7, 7)
    # odd block sizes over a 2 x 3 grid, like the reference's debug runs (test.cpp:4444-4445)
    seen = np.zeros((7, 7), int)
    for pr in range(2):
        for pc in range(3):
            loc = orc.block_cyclic_local(G, (3, 5), (2, 3), (pr, pc))
            for lr in range(loc.shape[0]):
                for lc in range(loc.shape[1]):
                    gr = orc.local_to_global(lr, 3, pr, 2)
                    gc = orc.local_to_global(lc, 5, pc, 3)
                    assert loc[lr, lc] == G[gr, gc]
                    seen[gr, gc] += 1
    assert (seen == 1).all()
