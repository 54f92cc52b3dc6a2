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


def _groups(*opts):
    r = subprocess.run([EXE, "groups", "--make-grm", "--bfile", os.path.join(PARTS, "in")] + list(opts),
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    out = {}
    for line in r.stdout.splitlines():
        if line.startswith("GROUP "):
            name, snps = line[6:].split(":", 1)
            out[name] = snps.split()
    return out, r.stdout


def test_group_file_semantics(tmp_path):
    """Genotype::groupSNPsByGene (genotype.cpp:1403-1446): `snp group [group ...]` lines, unknown SNPs skipped, a
    SNP may sit in several groups, groups come out in name order with their SNPs in set (lexicographic) order."""
    f = tmp_path / "g.txt"
    with open(f, "w") as h:
        for k in range(160):
            h.write(f"rs{k + 1} {'chrB' if k >= 100 else 'chrA'}{' extra' if k % 40 == 0 else ''}\n")
        h.write("rs_unknown chrA\n")
    got, log = _groups("--groups", str(f))
    assert list(got) == ["chrA", "chrB", "extra"] and "Created 3 SNP group(s)" in log
    assert got["chrA"] == sorted(f"rs{k + 1}" for k in range(100))
    assert got["chrB"] == sorted(f"rs{k + 1}" for k in range(100, 160))
    assert got["extra"] == sorted(["rs1", "rs41", "rs81", "rs121"])


def test_position_regions_and_overlap():
    """Genotype::groupSNPsByPosition (genotype.cpp:1347-1401): region j of a chromosome is
    [j*step, j*step + size) with step = size - overlap; names `<chr>_<start>-<end>`.  The fixture's .bim has
    chromosome 1 and positions 1..160, far below 1 kb: one region, and with --region-overlap a SNP belongs to every
    region containing it."""
    got, _ = _groups("--region-size", "1")
    assert list(got) == ["1_0-1000"] and len(got["1_0-1000"]) == 160
    got, _ = _groups("--region-size", "2", "--region-overlap", "1")
    assert list(got) == ["1_0-2000"] and len(got["1_0-2000"]) == 160


def test_group_all_and_min_snps(tmp_path):
    got, _ = _groups("--group-all")
    assert list(got) == ["all"] and len(got["all"]) == 160
    f = tmp_path / "g.txt"
    with open(f, "w") as h:
        for k in range(160):
            h.write(f"rs{k + 1} {'small' if k < 3 else 'big'}\n")
    got, log = _groups("--groups", str(f), "--min-snps-region", "5", "--out", str(tmp_path / "o"))
    assert list(got) == ["big"] and "1 regions with less than 5 SNPs have been filtered" in log
    assert open(str(tmp_path / "o") + ".filteredregions").read().startswith("small 3")


def test_regional_option_conflicts():
    r = subprocess.run([EXE, "groups", "--make-grm", "--bfile", os.path.join(PARTS, "in"), "--groups", "x",
                        "--region-size", "10"], capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "Only one type of regional analysis is allowed" in (r.stdout + r.stderr)


def test_position_regions_multi_region(tmp_path):
    """Several overlapping regions on two chromosomes against a direct evaluation of genotype.cpp:1361-1381:
    last region starting at or before pos = pos / step; first region ending after pos = 0 if pos < size else
    (pos - size) / step + 1."""
    import shutil
    bf = str(tmp_path / "pos")
    shutil.copy(os.path.join(PARTS, "in.fam"), bf + ".fam")
    snps = []
    with open(bf + ".bim", "w") as h:
        for k in range(200):
            chrom, pos = ("1", 137 * k + 5) if k < 120 else ("2", 211 * (k - 120) + 40000)
            snps.append((chrom, f"s{k}", pos))
            h.write(f"{chrom}\ts{k}\t0\t{pos}\tA\tG\n")
    size, overlap = 6000, 2000
    step = size - overlap
    want = {}
    for chrom, name, pos in snps:
        first = 0 if pos < size else (pos - size) // step + 1
        for j in range(first, pos // step + 1):
            want.setdefault(f"{chrom}_{j * step}-{j * step + size}", []).append(name)
    r = subprocess.run([EXE, "groups", "--make-grm", "--bfile", bf, "--region-size", "6", "--region-overlap", "2"],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    got = {}
    for line in r.stdout.splitlines():
        if line.startswith("GROUP "):
            name, s = line[6:].split(":", 1)
            got[name] = s.split()
    assert sorted(got) == sorted(want) and len(want) > 6
    for g in want:
        assert got[g] == sorted(want[g]), g


def test_host_add_equals_the_references_own_join(tmp_path):
    """tests/golden/parts/joined.* was written by the reference through its own Kernel::addKernels
    (`--bfile-list` + `--grm-join-method 0`, auxiliar.cpp:620-672: one unnormalised kernel per genotype file, added,
    then normalised).  The host mirror's addKernels on the two part files must reproduce it: N exactly, A to the
    rounding of the denormalise step it needs (the parts are read back normalised)."""
    out = str(tmp_path / "sum")
    run("add", os.path.join(PARTS, "a"), os.path.join(PARTS, "b"), out)
    A, N = load(out)
    AJ, NJ = load(os.path.join(PARTS, "joined"))
    AA, NA = load(os.path.join(PARTS, "all"))
    low = lower_mask(48)
    assert np.array_equal(N[low], NJ[low]) and np.array_equal(NJ[low], NA[low])
    assert np.max(np.abs(A[low] - AJ[low])) <= 1e-13
    assert np.max(np.abs(AJ[low] - AA[low])) <= 1e-12     # the reference's join vs its single-pass GRM
    assert open(os.path.join(PARTS, "joined.grm.ids")).read() == open(out + ".grm.ids").read()
    assert sorted(open(os.path.join(PARTS, "joined.grm.snps")).read().split()) == sorted(open(out + ".grm.snps").read().split())


@pytest.mark.parametrize("op,x,y,ref_name", [("sub", "all", "a", "loco_a"), ("sub", "all", "b", "loco_b"),
                                             ("add", "a", "b", "sum_ab")])
def test_host_algebra_reproduces_the_references_addKernels_files(op, x, y, ref_name, tmp_path):
    """tests/golden/parts/{loco_a,loco_b,sum_ab}.* were written by the reference's own Kernel::addKernels
    (kernel.cpp:1705-1859; LOCO: singlereml.cpp:298) on the reference's own part files, through oracle/_ref/grmalg.
    The host mirror performs the same operations in the same order: identical files."""
    out = str(tmp_path / "r")
    run(op, os.path.join(PARTS, x), os.path.join(PARTS, y), out)
    for ext in (".grm.ids", ".grm.snps", ".grm.dat"):
        assert open(out + ext, "rb").read() == open(os.path.join(PARTS, ref_name + ext), "rb").read(), ext
