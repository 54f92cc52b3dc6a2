"""GPU: the C++ host mirror (dissect_b200/host: Options/Genotype/Kernel/Matrix with the reference's names) driven
through its `--make-grm` CLI, compared file by file with what the reference's own `dissect --make-grm` wrote for
the same inputs (tests/golden)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import grm_oracle as orc
from tests.helpers import GOLDEN, load_case, lower_mask

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "dissect_b200", "host", "dgrm_make_grm")


def run_cli(case_dir, out, extra=()):
    cmd = [CLI, "--make-grm", "--bfile", os.path.join(case_dir, "in"), "--out", out]
    if os.path.exists(os.path.join(case_dir, "in.extract")):
        cmd += ["--extract", os.path.join(case_dir, "in.extract")]
    if os.path.exists(os.path.join(case_dir, "in.keep")):
        cmd += ["--keep", os.path.join(case_dir, "in.keep")]
    cmd += list(extra)
    return subprocess.run(cmd, capture_output=True, text=True, timeout=300)


@pytest.mark.parametrize("name", ["tiny7_m40", "n37_m64", "n130_m300_keep_extract", "n258_m500_flat",
                                  "n64_m200_nomiss", "sanitize_remove"])
def test_cli_files_match_reference(name, tmp_path):
    assert os.path.exists(CLI), "host CLI not built: python -c 'import __graft_entry__ as g; g.build()'"
    c = load_case(name)
    out = str(tmp_path / "gpu")
    r = run_cli(c["dir"], out, ["--grm-flat-norm"] if c["flat"] else [])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    for ext in (".grm.ids", ".grm.snps"):
        assert open(out + ext).read() == open(os.path.join(c["dir"], "ref" + ext)).read(), ext
    blob = open(out + ".grm.dat", "rb").read()
    assert len(blob) == len(c["ref_dat"]) and blob[:14] == c["ref_dat"][:14]
    n = len(c["ref_ids"])
    A, N = orc.unpack_grm_dat(blob, n)
    low = lower_mask(n)
    assert np.array_equal(N[low], c["ref_N"][low])
    m = len(open(out + ".grm.snps").read().split())
    assert np.max(np.abs(A[low] - c["ref_A"][low])) <= 3e-4 * max(1.0, (1000.0 / m) ** 0.5) * 2


def test_cli_sanitize_reject(tmp_path):
    c = load_case("sanitize_reject")
    out = str(tmp_path / "gpu")
    r = run_cli(c["dir"], out)
    assert r.returncode == 0
    assert c["meta"]["log_line"] in r.stdout            # same message as the reference's log
    assert not os.path.exists(out + ".grm.dat")         # and, like the reference, no GRM is written


def test_cli_badsnps(tmp_path):
    c = load_case("badsnp")
    out = str(tmp_path / "gpu")
    r = run_cli(c["dir"], out)
    assert r.returncode != 0
    assert "same genotype for all samples" in (r.stdout + r.stderr)
    assert open(out + ".badsnps").read() == open(os.path.join(c["dir"], "ref.badsnps")).read()
    assert not os.path.exists(out + ".grm.dat")


PARTS = os.path.join(GOLDEN, "parts")


def _check_group_files(prefix, ref_name, n=48):
    """<prefix>.grm.* against the files the reference wrote for the same SNP subset (tests/golden/parts)."""
    ref = os.path.join(PARTS, ref_name)
    assert open(prefix + ".grm.ids").read() == open(ref + ".grm.ids").read()
    assert open(prefix + ".grm.snps").read() == open(ref + ".grm.snps").read()
    blob, rblob = open(prefix + ".grm.dat", "rb").read(), open(ref + ".grm.dat", "rb").read()
    assert len(blob) == len(rblob) and blob[:14] == rblob[:14]
    A, N = orc.unpack_grm_dat(blob, n)
    A0, N0 = orc.unpack_grm_dat(rblob, n)
    low = lower_mask(n)
    assert np.array_equal(N[low], N0[low])
    m = len(open(prefix + ".grm.snps").read().split())
    assert np.max(np.abs(A[low] - A0[low])) <= 3e-4 * max(1.0, (1000.0 / m) ** 0.5) * 2


def test_cli_regional_groups_and_loco(tmp_path):
    """`--make-grm --groups F` (Analysis::makeGRM regional loop, analysis.cpp:67-110): one GRM per group from
    one upload of the .bed, plus --gpu-loco = global minus regional on the device (singlereml.cpp:293-298).
    Reference anchors: per-group `--extract` runs of the reference binary (tests/golden/parts)."""
    groups = tmp_path / "groups.txt"
    with open(groups, "w") as f:
        for k in range(160):
            f.write(f"rs{k + 1} {'a' if k < 70 else 'b'}\n")
        f.write("rs_not_in_bim a\n")                       # skipped like genotype.cpp:1424-1428
    out = str(tmp_path / "reg")
    r = run_cli(PARTS, out, ["--groups", str(groups), "--gpu-loco"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "Created 2 SNP group(s)" in r.stdout
    _check_group_files(out + ".a", "a")
    _check_group_files(out + ".b", "b")
    _check_group_files(out + ".loco.a", "b")              # all SNPs minus group a == group b
    _check_group_files(out + ".loco.b", "a")
    # ... and the files the reference's own Kernel::addKernels(1., all, -1., group) wrote (oracle/_ref/grmalg)
    _check_group_files(out + ".loco.a", "loco_a")
    _check_group_files(out + ".loco.b", "loco_b")
    # getGroup's side file (genotype.cpp:1583-1590): the group's SNPs in set order
    listed = open(out + ".a.region.a").read().split()
    assert listed == sorted(f"rs{k + 1}" for k in range(70))


def test_cli_group_all_and_min_snps(tmp_path):
    out = str(tmp_path / "all")
    r = run_cli(PARTS, out, ["--group-all"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_group_files(out + ".all", "all")
    groups = tmp_path / "g2.txt"
    with open(groups, "w") as f:
        for k in range(160):
            f.write(f"rs{k + 1} {'big' if k >= 5 else 'small'}\n")
    out2 = str(tmp_path / "min")
    r = run_cli(PARTS, out2, ["--groups", str(groups), "--min-snps-region", "10"])
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert os.path.exists(out2 + ".big.grm.dat") and not os.path.exists(out2 + ".small.grm.dat")
    assert open(out2 + ".filteredregions").read().startswith("small 5")


def test_cli_region_size_names(tmp_path):
    """--region-size groups by position (genotype.cpp:1347-1401); the .bim of the fixture has positions 1..160."""
    out = str(tmp_path / "pos")
    r = run_cli(PARTS, out, ["--region-size", "1"])       # 1 kb: every SNP falls in region 0 of chromosome 1
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    _check_group_files(out + ".1_0-1000", "all")


@pytest.mark.parametrize("engines", [2, 3])
def test_cli_multi_engine_files_identical(engines, tmp_path):
    """--gpu-engines G: the triangle split over G engines driven by one host rank (on a 1-GPU box they share the
    device), every engine writing its own runs of the .grm.dat: the files are byte-identical to the one-engine run,
    including the individuals sanitizeKernel removes (40 % missingness on two individuals)."""
    from dissect_b200 import bedio
    n, m = 2300, 420                       # n_pad = 2304: two raster row groups, 45 supertiles
    pk = bedio.synth_packed(n, m, 91, maf_lo=0.1, missing=0.02)
    rng = np.random.default_rng(5)
    for ind in (7, 1500):                  # two individuals with almost nothing typed: sanitizeKernel drops them
        for k in range(m):
            if rng.random() < 0.99:
                byte, sh = ind // 4, 2 * (ind % 4)
                pk[k, byte] = (int(pk[k, byte]) & (0xFF ^ (3 << sh))) | (0b01 << sh)
    bf = str(tmp_path / "in")
    bedio.write_bed(bf, pk, n)
    outs = {}
    for g in (1, engines):
        out = str(tmp_path / f"g{g}")
        r = subprocess.run([CLI, "--make-grm", "--bfile", bf, "--out", out, "--gpu-engines", str(g), "--gpu-chunk-snps", "256"],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[g] = (out, r.stdout)
    assert "individuals will be removed" in outs[1][1] and "individuals will be removed" in outs[engines][1]
    for ext in (".grm.ids", ".grm.snps", ".grm.dat"):
        assert open(outs[1][0] + ext, "rb").read() == open(outs[engines][0] + ext, "rb").read(), ext
    assert len(open(outs[1][0] + ".grm.ids").read().splitlines()) == n - 2


def test_cli_lazy_and_eager_host_matrices_agree(tmp_path):
    c = load_case("n130_m300_keep_extract")
    a, b = str(tmp_path / "lazy"), str(tmp_path / "eager")
    assert run_cli(c["dir"], a).returncode == 0
    assert run_cli(c["dir"], b, ["--gpu-eager-host"]).returncode == 0
    assert open(a + ".grm.dat", "rb").read() == open(b + ".grm.dat", "rb").read()


@pytest.mark.parametrize("method", [0, 1])
def test_cli_bfile_list_matches_reference_join(method, tmp_path):
    """--bfile-list + --grm-join-method (loadGRMUsingOptions, auxiliar.cpp:615-690): two genotype files with the same
    individuals and disjoint SNPs against the file the reference itself wrote through `--bfile-list ... --grm-join-method 0`
    (per-file kernels joined with Kernel::addKernels; tests/golden/parts/joined.*)."""
    from dissect_b200 import bedio
    d = os.path.join(GOLDEN, "parts")
    n, split = 48, 70
    pk = bedio.read_bed(os.path.join(d, "in"), n)
    bim = open(os.path.join(d, "in.bim")).read().splitlines()
    fam = open(os.path.join(d, "in.fam")).read()
    for name, (a, b) in (("pa", (0, split)), ("pb", (split, pk.shape[0]))):
        sub = str(tmp_path / name)
        bedio.write_bed(sub, pk[a:b], n)
        open(sub + ".fam", "w").write(fam)
        open(sub + ".bim", "w").write("\n".join(bim[a:b]) + "\n")
    lst = str(tmp_path / "list.txt")
    open(lst, "w").write(str(tmp_path / "pa") + "\n" + str(tmp_path / "pb") + "\n")
    out = str(tmp_path / "joined")
    r = subprocess.run([CLI, "--make-grm", "--bfile-list", lst, "--grm-join-method", str(method), "--out", out],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert f"Method {method}:" in r.stdout
    for ext in (".grm.ids", ".grm.snps"):
        assert open(out + ext).read() == open(os.path.join(d, "joined" + ext)).read(), ext
    ref_A, ref_N = orc.unpack_grm_dat(open(os.path.join(d, "joined.grm.dat"), "rb").read(), n)
    A, N = orc.unpack_grm_dat(open(out + ".grm.dat", "rb").read(), n)
    low = lower_mask(n)
    assert np.array_equal(N[low], ref_N[low])
    assert np.max(np.abs(A[low] - ref_A[low])) <= 3e-4 * (1000.0 / 160) ** 0.5
    # the same SNP in two files is an error, with the reference's message (auxiliar.cpp:651-655)
    open(lst, "w").write(str(tmp_path / "pa") + "\n" + str(tmp_path / "pa2") + "\n")
    for ext in (".bed", ".bim", ".fam"):
        open(str(tmp_path / "pa2") + ext, "wb").write(open(str(tmp_path / "pa") + ext, "rb").read())
    r = subprocess.run([CLI, "--make-grm", "--bfile-list", lst, "--out", out], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "There are SNPs repeated in different files" in r.stderr
