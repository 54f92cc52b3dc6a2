"""GPU: the resident cohort and the (S, N) algebra behind regional / LOCO GRMs,
and the standardised-genotype hand-back (§8b `dgrm_fetch_standardised`), through the C ABI.

Anchors: tests/golden/parts holds GRMs the *reference binary* wrote for SNP groups a = rs1..rs70, b = rs71..rs160,
for all 160 SNPs of one .bed, and for a + b joined by the reference's own Kernel::addKernels (`--bfile-list`,
`--grm-join-method 0`) (oracle/make_golden.py; the group GRMs are per-group `--extract` runs, because the reference's own
`--make-grm --groups` path aborts in normalizeGenotypes at this version: genotype.cpp:908-911 on a Genotype built by
the default constructor).  Group GRMs must match those files; saved - current must reproduce the other group
(Kernel::addKernels, kernel.cpp:1705-1859; LOCO, singlereml.cpp:298): N exactly, A within the operand tolerance."""
import os

import numpy as np
import pytest

from dissect_b200 import DgrmError, GrmEngine, bedio
from oracle import grm_oracle as orc
from tests.helpers import GOLDEN, lower_mask
from tests.test_gpu_parity import tol_for

pytestmark = pytest.mark.gpu

PARTS = os.path.join(GOLDEN, "parts")
N_FAM, M_FILE, SPLIT = 48, 160, 70


def ref(prefix):
    return orc.unpack_grm_dat(open(os.path.join(PARTS, prefix + ".grm.dat"), "rb").read(), N_FAM)


def fetch(eng, which):
    n = eng.num_individuals
    return eng.fetch_block_cyclic(which, bs=(n, n))


def group_mask(a, b):
    k = np.zeros(M_FILE, np.uint8)
    k[a:b] = 1
    return k


@pytest.fixture()
def eng():
    e = GrmEngine(chunk_snps=256)
    e.set_individuals(N_FAM)
    e.store_bed_file(os.path.join(PARTS, "in.bed"), M_FILE)
    assert e.stored_snps == M_FILE
    yield e
    e.close()


def build_group(e, a, b):
    e.begin(b - a)
    e.push_stored(row_keep=group_mask(a, b))
    e.finish()


def test_group_grms_from_resident_rows_match_reference_files(eng):
    low = lower_mask(N_FAM)
    for name, (a, b) in {"a": (0, SPLIT), "b": (SPLIT, M_FILE), "all": (0, M_FILE)}.items():
        build_group(eng, a, b)
        A0, N0 = ref(name)
        assert np.array_equal(fetch(eng, "N")[low], N0[low]), name
        assert np.max(np.abs(fetch(eng, "A")[low] - A0[low])) <= tol_for("fp16", b - a), name
    # a contiguous range pushed without a mask is the same GRM
    eng.begin(M_FILE - SPLIT)
    eng.push_stored(first=SPLIT, nrows=M_FILE - SPLIT)
    eng.finish()
    A0, N0 = ref("b")
    assert np.array_equal(fetch(eng, "N")[low], N0[low])


def test_loco_view_is_global_minus_regional(eng, tmp_path):
    low = lower_mask(N_FAM)
    build_group(eng, 0, M_FILE)
    S_all = fetch(eng, "S").copy()
    eng.save("copy")
    build_group(eng, 0, SPLIT)                       # regional GRM (group a)
    S_a = fetch(eng, "S").copy()
    eng.set_view("saved-current")                    # addKernels(1., global, -1., regional)
    A_b, N_b = ref("b")
    assert np.array_equal(fetch(eng, "N")[low], N_b[low])
    assert np.max(np.abs(fetch(eng, "A")[low] - A_b[low])) <= tol_for("fp16", M_FILE - SPLIT)
    assert np.array_equal(fetch(eng, "S")[low], (S_all - S_a)[low])      # exact fp64 subtraction of the stores
    A_l, N_l = ref("loco_a")                          # the reference's own addKernels(1., all, -1., a) on its files
    assert np.array_equal(fetch(eng, "N")[low], N_l[low])
    assert np.max(np.abs(fetch(eng, "A")[low] - A_l[low])) <= tol_for("fp16", M_FILE - SPLIT)
    # every reader sees the view: the file writer and the low-overlap scan
    path = str(tmp_path / "loco.grm.dat")
    eng.write_grm_dat(path)
    A2, N2 = orc.unpack_grm_dat(open(path, "rb").read(), N_FAM)
    assert np.array_equal(N2[low], N_b[low]) and np.array_equal(A2[low], fetch(eng, "A")[low])
    thr = float(np.sort(N_b[low])[40])
    pairs, cnt = eng.low_overlap_pairs(min_overlap=thr + 0.1)
    got = set(map(tuple, pairs.tolist()))
    want = {(i, j) for i in range(N_FAM) for j in range(i) if -0.1 < N_b[i, j] < thr}
    assert got == want and cnt == len(want)
    # the saved GRM itself, and back to the current one
    eng.set_view("saved")
    A_all, N_all = ref("all")
    assert np.array_equal(fetch(eng, "N")[low], N_all[low])
    assert np.array_equal(fetch(eng, "S")[low], S_all[low])
    eng.set_view("current")
    assert np.array_equal(fetch(eng, "S")[low], S_a[low])


def test_sum_of_groups_accumulates_to_the_global_grm(eng):
    """The reference's test20 invariant (test.cpp:1722-1729): N exactly, A to accumulation order."""
    low = lower_mask(N_FAM)
    build_group(eng, 0, SPLIT)
    eng.save("copy")
    build_group(eng, SPLIT, M_FILE)
    eng.set_view("saved+current")
    A_all, N_all = ref("all")
    assert np.array_equal(fetch(eng, "N")[low], N_all[low])
    assert np.max(np.abs(fetch(eng, "A")[low] - A_all[low])) <= tol_for("fp16", M_FILE)
    eng.save("add")                                   # saved = a + b
    eng.set_view("saved")
    assert np.array_equal(fetch(eng, "N")[low], N_all[low])
    assert np.max(np.abs(fetch(eng, "A")[low] - A_all[low])) <= tol_for("fp16", M_FILE)
    # ... and the file the reference wrote through its own Kernel::addKernels (--bfile-list, --grm-join-method 0)
    A_j, N_j = ref("joined")
    assert np.array_equal(fetch(eng, "N")[low], N_j[low])
    assert np.max(np.abs(fetch(eng, "A")[low] - A_j[low])) <= tol_for("fp16", M_FILE)
    eng.save("drop")
    with pytest.raises(DgrmError):
        eng.set_view("saved")


def test_streamed_outputs_honour_the_view(eng):
    """Bound (streaming) outputs filled during dgrm_finish carry the LOCO values directly."""
    from dissect_b200.engine import PinnedArray
    low = lower_mask(N_FAM)
    build_group(eng, 0, M_FILE)
    eng.save("copy")
    eng.set_view("saved-current")
    A = PinnedArray((N_FAM, N_FAM), np.float64, order="F")
    Nn = PinnedArray((N_FAM, N_FAM), np.float64, order="F")
    eng.bind_output("A", A.ptr, N_FAM)
    eng.bind_output("N", Nn.ptr, N_FAM)
    build_group(eng, 0, SPLIT)
    A_b, N_b = ref("b")
    assert np.array_equal(Nn.array[low], N_b[low])
    assert np.array_equal(A.array[low], fetch(eng, "A")[low])
    eng.unbind_outputs()


@pytest.mark.parametrize("bs,grid", [((256, 256), (1, 1)), ((16, 8), (2, 2)), ((7, 5), (3, 2))])
def test_standardised_genotypes_zero_ulp(bs, grid):
    """genotype->genotypes after normalizeGenotypes (genotype.cpp:888-961), with --keep and a SNP subset."""
    n_fam, m = 131, 300
    pk = bedio.synth_packed(n_fam, m, 31, maf_lo=0.05, missing=0.04)
    keep = (np.arange(n_fam) % 5 != 2).astype(np.int32)
    rk = (np.arange(m) % 3 != 1).astype(np.uint8)
    codes = orc.decode_rows(pk[rk != 0], n_fam, keep)
    st = orc.snp_stats(codes)
    Z = orc.standardise(codes, st["p2"], st["sd"])              # [m_kept, n_kept]
    e = GrmEngine(chunk_snps=256)
    e.set_individuals(n_fam, keep)
    e.store_bed_host(np.ascontiguousarray(pk))
    e.begin(int(rk.sum()))
    e.push_stored(row_keep=rk)
    e.finish()
    for r in range(grid[0]):
        for c in range(grid[1]):
            got = e.fetch_standardised(bs=bs, grid=grid, coords=(r, c))
            rows = [g for g in range(Z.shape[0]) if (g // bs[0]) % grid[0] == r]
            cols = [g for g in range(Z.shape[1]) if (g // bs[1]) % grid[1] == c]
            want = Z[np.ix_(rows, cols)]
            assert got.shape == want.shape
            assert np.array_equal(got.view(np.uint64), np.ascontiguousarray(want).view(np.uint64).reshape(want.shape))
    e.close()


def test_standardised_needs_resident_rows():
    pk = bedio.synth_packed(40, 64, 5, missing=0.02)
    e = GrmEngine(chunk_snps=256)
    e.set_individuals(40)
    e.begin(64)
    e.push_bed_host(np.ascontiguousarray(pk))
    e.finish()
    with pytest.raises(DgrmError):
        e.fetch_standardised()
    e.close()


@pytest.mark.parametrize("name", ["tiny7_m40", "n37_m64"])
def test_standardised_genotypes_equal_the_references_own_matrix(name):
    """dgrm_fetch_standardised against ref.z.bin = genotype->genotypes as the *reference binary* left it after
    normalizeGenotypes (dumped by oracle/snpdump.cpp): bit for bit."""
    from tests.helpers import load_case
    c = load_case(name)
    zfile = os.path.join(c["dir"], "ref.z.bin")
    if not os.path.exists(zfile):
        pytest.skip("no ref.z.bin for this case")
    e = GrmEngine(chunk_snps=256)
    e.set_individuals(c["n_fam"], c["keep_mask"])
    e.store_bed_host(np.ascontiguousarray(c["packed"]))
    m = c["m_file"] if c["snp_keep"] is None else int(c["snp_keep"].sum())
    e.begin(m)
    e.push_stored(row_keep=c["snp_keep"])
    e.finish()
    n = e.num_individuals
    z_ref = np.fromfile(zfile).reshape(2, m, n)[0]
    got = e.fetch_standardised(bs=(m, n))
    assert np.array_equal(np.ascontiguousarray(got).view(np.uint64), z_ref.view(np.uint64))
    e.close()
