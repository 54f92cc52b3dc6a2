"""CPU: the numpy restatement of the GWAS LD step (oracle/grm_oracle.ld) against what the reference's own objects
produced (tests/golden/ld_*, written by oracle/_ref/ldcorr via oracle/make_golden_ld.py):
correlations = G G^T o 1/(M M^T) over the lower triangle and the index lists of getGlobalIndexOutsideRange."""
import os

import numpy as np
import pytest

from tests.helpers import load_ld_case
from oracle import grm_oracle as orc

LD_CASES = ["ld_n61_m40", "ld_n130_m90_keep_extract", "ld_n300_m270"]


@pytest.mark.parametrize("name", LD_CASES)
def test_ld_restatement_matches_reference(name):
    c = load_ld_case(name)
    codes = orc.decode_rows(c["rows"], c["n_fam"], c["keep_mask"])
    got = orc.ld(codes, c["threshold"])
    m = codes.shape[0]
    assert c["ref_C"].shape == (m, m)
    low = np.tril(np.ones((m, m), bool))
    # S: two correct fp64 dsyrk differ only in summation order; N is exact, so C agrees to ~1e-15
    assert np.max(np.abs(got["C"][low] - c["ref_C"][low])) < 1e-13
    # index lists: the reference reports diagonal hits once and off-diagonal hits in both orders (matrix.cpp:4287-4299)
    ref_strict = sorted({(r, cc) for r, cc in c["ref_pairs"] if r > cc})
    assert got["pairs"] == ref_strict and len(ref_strict) > 0
    both = {(r, cc) for r, cc in c["ref_pairs"]}
    assert all((cc, r) in both for r, cc in ref_strict)
    diag_hits = sorted(r for r, cc in c["ref_pairs"] if r == cc)
    want_diag = [k for k in range(m) if got["C"][k, k] > c["threshold"] or got["C"][k, k] < -c["threshold"]]
    assert diag_hits == want_diag
