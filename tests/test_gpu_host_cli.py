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
