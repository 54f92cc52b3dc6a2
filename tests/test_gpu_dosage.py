"""GPU parity tests of the dosage ingest (dgrm_push_dosage_host): the CUDA path through the C ABI
against the oracle's restatement of the reference's BGEN standardisation (genotypebgen.cpp:239-282, 723-779) and GRM
(kernel.cpp:93-100).  Bars: N bit-exact; mean / std / p1 / p2 0 ulp (the device accumulates each row in the
reference's sample order); A within the operand tolerance of tests/test_gpu_parity.py; the exact fp64 diagonal to 1e-9.  PARITY UNPINNED (see oracle/grm_oracle.py::dosage_standardise)."""
import numpy as np
import pytest

from dissect_b200 import DgrmError, GrmEngine, bedio, bgenio
from oracle import grm_oracle as orc
from tests.helpers import lower_mask
from tests.test_dosage_cpu import synth_probs
from tests.test_gpu_parity import fetch_all, tol_for

pytestmark = pytest.mark.gpu


def synth_dosage(m, n, seed, missing=0.03):
    p_aa, p_ab, miss = synth_probs(m, n, seed, 255, missing)
    a, b = p_aa / 255.0, p_ab / 255.0
    return b + 2.0 * (1.0 - (a + b)), miss


def check_stats(eng, st, m):
    """mean / std are accumulated in the reference's sample order on the device: 0 ulp."""
    got = eng.snp_stats()
    assert np.all(got["nn"][:m] == -1) and not got["freq"][:m].any()     # genotypebgen.cpp:733
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(got[k][:m].view(np.uint64), st[k].view(np.uint64)), k


def compare(eng, ref, operand, m, n):
    got = fetch_all(eng)
    low = lower_mask(n)
    assert np.array_equal(got["N"][low], ref["N"][low])
    ok = low & (ref["N"] > 0)
    assert np.max(np.abs(got["A"][ok] - ref["A"][ok])) <= tol_for(operand, m)
    assert np.max(np.abs(np.diag(got["S"]) - np.diag(ref["S"]))) <= 1e-9
    return got


@pytest.mark.parametrize("operand,count,cg", [("fp16", "fp4", 2), ("fp16", "fp4", 1), ("fp16", "mma", 2),
                                              ("bf16", "fp4", 2), ("fp16x2", "mma", 2), ("fp16", "popc", 1)])
def test_dosage_grm_vs_oracle(operand, count, cg):
    n, m = 515, 1900
    d, miss = synth_dosage(m, n, 5)
    d[7] = 1.25                      # no variance (and the running sum is exact): dropped from S and N
    miss[11] = True                  # nobody typed: dropped as well
    ref = orc.dosage_grm(d, ~miss)
    eng = GrmEngine(operand=operand, count_method=count, cta_group=cg, chunk_snps=512)
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_dosage_host(d, miss)
    eng.finish()
    assert eng.num_snps == m
    check_stats(eng, ref["stats"], m)
    compare(eng, ref, operand, m, n)
    eng.close()


def test_row_at_a_time_equals_block_pushes():
    """A variant parser delivers one row per call: the staging makes that bit-identical to one large push."""
    n, m = 300, 700
    d, miss = synth_dosage(m, n, 9)
    outs = []
    for mode in ("block", "rows", "ragged"):
        eng = GrmEngine(chunk_snps=256)
        eng.set_individuals(n)
        eng.begin(m)
        if mode == "block":
            eng.push_dosage_host(d, miss)
        elif mode == "rows":
            for k in range(m):
                eng.push_dosage_host(d[k:k + 1], miss[k:k + 1] if miss[k].any() else None)
        else:
            for a, b in ((0, 1), (1, 255), (255, 513), (513, 700)):
                eng.push_dosage_host(d[a:b], miss[a:b])
        eng.finish()
        outs.append((fetch_all(eng), eng.snp_stats()))
        eng.close()
    low = lower_mask(n)
    for got, st in outs[1:]:
        for k in ("S", "N", "A"):
            assert np.array_equal(got[k][low], outs[0][0][k][low]), k
        for k in ("p1", "p2", "sd"):
            assert np.array_equal(st[k], outs[0][1][k])
    ref = orc.dosage_grm(d, ~miss)
    assert np.array_equal(outs[0][0]["N"][low], ref["N"][low])


def test_staging_smaller_than_the_chunk(monkeypatch):
    """A device staging buffer that holds fewer rows than one chunk (and fewer than one pinned transfer buffer): several
    batches fill a chunk, a transfer buffer is split over two batches, the last batch ends on a partial k-block."""
    monkeypatch.setenv("DGRM_DOSAGE_STAGE_MB", "64")
    n, m = 2100, 18000
    rng = np.random.default_rng(4)
    f = rng.uniform(0.1, 0.5, size=(m, 1))
    d = rng.binomial(2, f, size=(m, n)).astype(np.float64) + rng.integers(-20, 21, size=(m, n)) / 255.0
    d = np.clip(d, 0.0, 2.0)
    miss = rng.random((m, n)) < 0.01
    ref = orc.dosage_grm(d, ~miss)
    eng = GrmEngine()                     # default chunk: 16384 SNPs
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_dosage_host(d[:5000], miss[:5000])
    eng.push_dosage_host(d[5000:], miss[5000:])
    eng.finish()
    check_stats(eng, ref["stats"], m)
    compare(eng, ref, "fp16", m, n)
    eng.close()


def test_no_missing_mask_and_flat_norm():
    n, m = 260, 600
    d, _ = synth_dosage(m, n, 13, missing=0.0)
    ref = orc.dosage_grm(d, None, flat_norm=True)
    eng = GrmEngine(flat_norm=True, chunk_snps=256)
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_dosage_host(d)
    eng.finish()
    got = compare(eng, ref, "fp16", m, n)
    assert np.all(got["N"][lower_mask(n)] == float(m))
    eng.close()
    # a zero-variance row changes the divisor under --grm-flat-norm: refused, with the way out in the message
    d[3] = 0.0
    eng = GrmEngine(flat_norm=True, chunk_snps=256)
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_dosage_host(d)
    with pytest.raises(DgrmError, match="zero variance"):
        eng.finish()
    eng.close()


def test_bed_and_dosage_rows_in_one_grm():
    """--bfile-list style SNP-wise join of a .bed source and a dosage source over the same individuals."""
    n, m1, m2 = 333, 450, 390
    pk = bedio.synth_packed(n, m1, 77, missing=0.02)
    codes = orc.decode_rows(pk, n)
    st = orc.snp_stats(codes)
    zb = orc.standardise(codes, st["p2"], st["sd"])
    qb = (codes != 0).astype(np.float64)
    d, miss = synth_dosage(m2, n, 78)
    for order in ("bed-first", "dosage-first", "interleaved"):
        eng = GrmEngine(chunk_snps=256)
        eng.set_individuals(n)
        eng.begin(m1 + m2)
        if order == "bed-first":
            eng.push_bed_host(pk)
            eng.push_dosage_host(d, miss)
        elif order == "dosage-first":
            eng.push_dosage_host(d, miss)
            eng.push_bed_host(pk)
        else:
            eng.push_bed_host(pk[:200])
            eng.push_dosage_host(d[:100], miss[:100])
            eng.push_bed_host(pk[200:])
            eng.push_dosage_host(d[100:], miss[100:])
        eng.finish()
        ref = orc.dosage_grm(d, ~miss, extra_z=zb, extra_q=qb)
        compare(eng, ref, "fp16", m1 + m2, n)
        got = eng.snp_stats()
        if order == "bed-first":
            assert np.array_equal(got["freq"][:m1], st["freq"]) and np.all(got["nn"][m1:] == -1)
            assert np.array_equal(got["sd"][:m1], st["sd"])
        eng.close()


@pytest.mark.parametrize("layout,bits", [(2, 8), (2, 16), (1, 16)])
def test_bgen_file_to_grm(tmp_path, layout, bits):
    m, n = 420, 271
    denom = 32768 if layout == 1 else (1 << bits) - 1
    p_aa, p_ab, miss = synth_probs(m, n, 31, denom)
    path = str(tmp_path / "c.bgen")
    bgenio.write_bgen(path, p_aa, p_ab, miss, layout=layout, bits=bits)
    keep = (np.arange(n) % 5 != 2).astype(np.int32)
    nk = int(keep.sum())
    blocks = list(bgenio.iter_dosage_blocks(path, keep_mask=keep, block=1 << 20))
    ref = orc.dosage_grm(blocks[0][1], blocks[0][2] == 0)
    eng = GrmEngine(chunk_snps=256)
    eng.set_individuals(nk)
    eng.begin(m)
    meta = bgenio.push_bgen(eng, path, keep_mask=keep, block=100)
    eng.finish()
    assert len(meta) == m and eng.num_individuals == nk
    check_stats(eng, ref["stats"], m)
    compare(eng, ref, "fp16", m, nk)
    eng.close()


def test_dosage_refused_in_ld_mode_and_over_capacity():
    eng = GrmEngine()
    eng.ld_begin(64, 32)
    with pytest.raises(DgrmError, match="LD mode"):
        eng.push_dosage_host(np.zeros((2, 64)))
    eng.close()
    eng = GrmEngine()
    eng.set_individuals(64)
    eng.begin(3)
    with pytest.raises(DgrmError, match="capacity"):
        eng.push_dosage_host(np.ones((4, 64)))
    with pytest.raises(DgrmError, match="bad rows/pitch"):
        eng.push_dosage_host(np.ones((2, 63)))
    eng.close()


@pytest.mark.parametrize("engines,host_decode", [(1, False), (3, False), (1, True)])
def test_cli_bgen_make_grm(tmp_path, engines, host_decode):
    """`dgrm_make_grm --bgen F --make-grm` (the reference's `dissect --bgen F --make-grm`, analysis.cpp:515-522): the
    C++ host mirror parses the container, the engines standardise and contract, the files follow the reference's
    conventions (ids = (id, id), zero-variance SNPs listed in <out>.snps.novariance and absent from .grm.snps)."""
    import os
    import subprocess
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "dissect_b200", "host", "dgrm_make_grm")
    assert os.path.exists(cli), "host CLI not built: python -c 'import __graft_entry__ as g; g.build()'"
    m, n = 610, 301
    p_aa, p_ab, miss = synth_probs(m, n, 91, 255)
    p_aa[17], p_ab[17], miss[17] = 255, 0, False         # P(AA) = 1 for everybody: dosage 0, no variance
    prefix = str(tmp_path / "cohort")
    ids = ["id%d" % i for i in range(n)]
    bgenio.write_bgen(prefix + ".bgen", p_aa, p_ab, miss, sample_ids=ids)
    keep = [i for i in range(n) if i % 7 != 3]
    with open(prefix + ".keep", "w") as f:
        for i in keep:
            f.write("id%d id%d\n" % (i, i))
    out = str(tmp_path / "out")
    r = subprocess.run([cli, "--make-grm", "--bgen", prefix, "--keep", prefix + ".keep", "--out", out, "--gpu-chunk-snps", "256",
                        "--gpu-engines", str(engines)] + (["--gpu-bgen-host-decode"] if host_decode else []),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "1 SNPs which have zero variance and will be removed" in r.stdout
    assert open(out + ".snps.novariance").read().split() == ["rs17"]
    assert open(out + ".grm.snps").read().split() == ["rs%d" % k for k in range(m) if k != 17]
    assert [l.split() for l in open(out + ".grm.ids")] == [["id%d" % i, "id%d" % i] for i in keep]
    mask = np.zeros(n, np.int32)
    mask[keep] = 1
    _, dos, mis = list(bgenio.iter_dosage_blocks(prefix + ".bgen", keep_mask=mask, block=1 << 20))[0]
    ref = orc.dosage_grm(dos, mis == 0)
    nk = len(keep)
    A, N = orc.unpack_grm_dat(open(out + ".grm.dat", "rb").read(), nk)
    low = lower_mask(nk)
    assert np.array_equal(N[low], ref["N"][low])
    assert np.max(np.abs(A[low] - ref["A"][low])) <= tol_for("fp16", m)


def test_constant_row_follows_the_reference_rounding():
    """A row with the same inexact dosage for everybody: the reference's running sum rounds, its mean misses the common
    value by an ulp or so and `standardDev != 0.` keeps the SNP (genotypebgen.cpp:746).  Whatever the reference's
    arithmetic decides, the device decides the same, because it accumulates in the same order."""
    n, m = 258, 300
    d, miss = synth_dosage(m, n, 17)
    consts = [(200, 30), (100, 100), (1, 253), (77, 13), (255, 0), (0, 255), (128, 0)]
    for k, (a, b) in enumerate(consts):
        d[3 * k] = b / 255.0 + 2.0 * (1.0 - (a / 255.0 + b / 255.0))
        miss[3 * k] = False
    ref = orc.dosage_grm(d, ~miss)
    sd = ref["stats"]["sd"][[3 * k for k in range(len(consts))]]
    assert (sd == 0).any() and (sd != 0).any()          # both outcomes occur among these rows
    eng = GrmEngine(chunk_snps=256)
    eng.set_individuals(n)
    eng.begin(m)
    eng.push_dosage_host(d, miss)
    eng.finish()
    check_stats(eng, ref["stats"], m)
    got = fetch_all(eng)
    assert np.array_equal(got["N"][lower_mask(n)], ref["N"][lower_mask(n)])
    eng.close()


@pytest.mark.parametrize("layout,bits", [(2, 8), (2, 16), (2, 11), (2, 3), (2, 32), (1, 16)])
def test_device_unpack_of_bgen_blocks_is_bit_identical(tmp_path, layout, bits):
    """dgrm_push_bgen_blocks_host (inflated blocks unpacked by bgen_unpack_kernel) against dgrm_push_dosage_host fed by
    the host reader: same statistics, S, N and A to the last bit, with a sample selection and missing samples."""
    m, n = 530, 297
    denom = 32768 if layout == 1 else (1 << bits) - 1
    p_aa, p_ab, miss = synth_probs(m, n, 50 + bits, denom)
    path = str(tmp_path / "u.bgen")
    bgenio.write_bgen(path, p_aa, p_ab, miss, layout=layout, bits=bits)
    keep = (np.arange(n) % 6 != 4).astype(np.int32)
    nk = int(keep.sum())
    outs = []
    for device_unpack in (True, False):
        eng = GrmEngine(chunk_snps=256)
        eng.set_individuals(nk)
        eng.begin(m)
        meta = bgenio.push_bgen(eng, path, keep_mask=keep, block=77, device_unpack=device_unpack)
        eng.finish()
        assert len(meta) == m
        outs.append((fetch_all(eng), eng.snp_stats()))
        eng.close()
    low = lower_mask(nk)
    for k in ("S", "N", "A"):
        assert np.array_equal(outs[0][0][k][low], outs[1][0][k][low]), k
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(outs[0][1][k].view(np.uint64), outs[1][1][k].view(np.uint64)), k
    blocks = list(bgenio.iter_dosage_blocks(path, keep_mask=keep, block=1 << 20))
    ref = orc.dosage_grm(blocks[0][1], blocks[0][2] == 0)
    assert np.array_equal(outs[0][0]["N"][low], ref["N"][low])
    for k in ("p1", "p2", "sd"):
        assert np.array_equal(outs[0][1][k][:m].view(np.uint64), ref["stats"][k].view(np.uint64)), k


def test_bgen_blocks_mixed_with_dosage_and_bed_rows(tmp_path):
    n, m = 200, 300
    p_aa, p_ab, miss = synth_probs(m, n, 61, 255)
    path = str(tmp_path / "x.bgen")
    bgenio.write_bgen(path, p_aa, p_ab, miss)
    _, dos, mis = list(bgenio.iter_dosage_blocks(path, block=1 << 20))[0]
    raw = list(bgenio.iter_raw_blocks(path, block=1 << 20))[0]
    data, off, hdr = raw[1], raw[2], raw[3]
    pk = bedio.synth_packed(n, 120, 5, missing=0.02)
    eng = GrmEngine(chunk_snps=256)
    eng.set_individuals(n)
    eng.begin(m + 120)
    eng.push_dosage_host(dos[:50], mis[:50])
    eng.push_bgen_blocks_host(data[off[50]:off[170]], off[50:171] - off[50], hdr.layout, hdr.n_samples)
    eng.push_bed_host(pk)
    eng.push_bgen_blocks_host(data[off[170]:off[260]], off[170:261] - off[170], hdr.layout, hdr.n_samples)
    eng.push_dosage_host(dos[260:], mis[260:])
    eng.finish()
    codes = orc.decode_rows(pk, n)
    st = orc.snp_stats(codes)
    zb = orc.standardise(codes, st["p2"], st["sd"])
    # SNP order: dosage[0:170], bed, dosage[170:]; the GRM does not depend on it
    ref = orc.dosage_grm(dos, mis == 0, extra_z=zb, extra_q=(codes != 0).astype(np.float64))
    compare(eng, ref, "fp16", m + 120, n)
    got = eng.snp_stats()
    rs = orc.dosage_standardise(dos, mis == 0)
    assert np.array_equal(got["sd"][:170], rs["sd"][:170]) and np.array_equal(got["sd"][170:290], st["sd"])
    assert np.array_equal(got["sd"][290:m + 120], rs["sd"][170:])
    eng.close()


def test_bgen_blocks_refusals(tmp_path):
    n, m = 40, 6
    p = np.full((m, n), 80, np.uint64)
    path = str(tmp_path / "r.bgen")
    bgenio.write_bgen(path, p, p, compressed=False)
    _, data, off, hdr = list(bgenio.iter_raw_blocks(path))[0]
    data = data.copy()
    eng = GrmEngine()
    eng.set_individuals(n)
    eng.begin(m)
    bad = data.copy()
    bad[off[2] + 8 + n] = 1                                  # phased flag of block 2
    with pytest.raises(DgrmError, match="unphased"):
        eng.push_bgen_blocks_host(bad, off, 2, n)
    bad = data.copy()
    bad[off[1] + 4] = 3                                      # three alleles
    with pytest.raises(DgrmError, match="non-biallelic"):
        eng.push_bgen_blocks_host(bad, off, 2, n)
    with pytest.raises(DgrmError, match="samples are kept"):
        eng.push_bgen_blocks_host(data, off, 2, n, sample_keep=np.arange(n) % 2)
    bad = data.copy()
    bad[off[3] + 8 + 5] = 3                                  # sample 5 of block 3 is triploid: found on the device
    eng.push_bgen_blocks_host(bad, off, 2, n)
    with pytest.raises(DgrmError, match="non-diploid"):
        eng.finish()
    eng.close()
