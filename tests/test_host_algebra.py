"""CPU: the host mirror's file and algebra paths (readGRM, writeGRM, addKernels — SURVEY §8f rows 1-2) on GRM files
written by the reference binary: the reference's test20 / test57 invariants (test.cpp:1722-1729, 4383-4391)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import grm_oracle as orc
from tests.helpers import GOLDEN, lower_mask

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "dissect_b200", "host", "selftest")
PARTS = os.path.join(GOLDEN, "parts")

pytestmark = pytest.mark.skipif(not os.path.exists(EXE), reason="host selftest not built (build())")


def run(*args):
    r = subprocess.run([EXE] + list(args), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]


def load(prefix, n=48):
    return orc.unpack_grm_dat(open(prefix + ".grm.dat", "rb").read(), n)


def test_read_write_round_trip_is_exact(tmp_path):
    out = str(tmp_path / "rt")
    run("roundtrip", os.path.join(PARTS, "all"), out)
    for ext in (".grm.dat", ".grm.ids", ".grm.snps"):
        assert open(out + ext, "rb").read() == open(os.path.join(PARTS, "all" + ext), "rb").read(), ext


def test_sum_of_parts_equals_merged(tmp_path):
    out = str(tmp_path / "sum")
    run("add", os.path.join(PARTS, "a"), os.path.join(PARTS, "b"), out)
    A, N = load(out)
    A0, N0 = load(os.path.join(PARTS, "all"))
    low = lower_mask(48)
    assert np.array_equal(N[low], N0[low])                       # test20: N exactly equal
    assert np.max(np.abs(A[low] - A0[low])) <= 1e-10             # test20: 1e-10
    # SNP bookkeeping: b's SNPs then a's (kernel.cpp:1826-1827)
    got = open(out + ".grm.snps").read().split()
    assert sorted(got) == sorted(open(os.path.join(PARTS, "all.grm.snps")).read().split()) and got[0] == "rs71"


def test_merged_minus_part_equals_other_part(tmp_path):
    """LOCO algebra (singlereml.cpp:293-298): global - regional on the (S, N) pair."""
    out = str(tmp_path / "diff")
    run("sub", os.path.join(PARTS, "all"), os.path.join(PARTS, "a"), out)
    A, N = load(out)
    B, NB = load(os.path.join(PARTS, "b"))
    low = lower_mask(48)
    assert np.array_equal(N[low], NB[low])
    assert np.max(np.abs(A[low] - B[low])) <= 1e-10
    assert open(out + ".grm.snps").read() == open(os.path.join(PARTS, "b.grm.snps")).read()


def test_overlapping_snp_sets_cannot_be_added(tmp_path):
    r = subprocess.run([EXE, "add", os.path.join(PARTS, "a"), os.path.join(PARTS, "all"), str(tmp_path / "x")],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 1 and "cannot be added together if the SNPs" in (r.stdout + r.stderr)
