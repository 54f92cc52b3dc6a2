"""CPU: fresh seeds through the reference's own binary (oracle/_ref, when it has been built in this
container) against the oracle restatement — the pin that makes the restatement trustworthy."""
import os

import numpy as np
import pytest

from dissect_b200 import bedio
from oracle import grm_oracle as orc
from oracle import ref_runner
from tests.helpers import lower_mask

pytestmark = pytest.mark.skipif(not ref_runner.available(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.mark.parametrize("n,m,missing,seed", [(61, 300, 0.03, 1), (258, 700, 0.01, 2), (400, 1500, 0.0, 3)])
def test_reference_binary_equals_restatement(tmp_path, n, m, missing, seed):
    pk = bedio.synth_packed(n, m, seed, missing=missing)
    bf = str(tmp_path / "c")
    bedio.write_bed(bf, pk, n)
    rc, log = ref_runner.run_make_grm(bf, str(tmp_path / "ref"))
    assert rc == 0, log[-2000:]
    st = ref_runner.run_snpdump(bf, str(tmp_path / "ref"))
    g = orc.grm(orc.decode_rows(pk, n))
    assert np.array_equal(st["freq"], g["stats"]["freq"]) and np.array_equal(st["nn"], g["stats"]["nn"])
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(st[k].view(np.uint64), g["stats"][k].view(np.uint64))
    A, N = orc.unpack_grm_dat(open(str(tmp_path / "ref.grm.dat"), "rb").read(), n)
    low = lower_mask(n)
    assert np.array_equal(N[low], g["N"][low])
    assert np.max(np.abs(A[low] - g["A"][low])) <= 1e-13
    assert open(str(tmp_path / "ref.grm.ids")).read() == "".join(f"f{i + 1} i{i + 1}\n" for i in range(n))


def test_reference_flat_norm(tmp_path):
    n, m = 80, 200
    pk = bedio.synth_packed(n, m, 9, missing=0.05)
    bf = str(tmp_path / "c")
    bedio.write_bed(bf, pk, n)
    rc, log = ref_runner.run_make_grm(bf, str(tmp_path / "ref"), flat_norm=True)
    assert rc == 0, log[-2000:]
    A, N = orc.unpack_grm_dat(open(str(tmp_path / "ref.grm.dat"), "rb").read(), n)
    g = orc.grm(orc.decode_rows(pk, n), flat_norm=True)
    low = lower_mask(n)
    assert np.all(N[low] == m) and np.max(np.abs(A[low] - g["A"][low])) <= 1e-13
