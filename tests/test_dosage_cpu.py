"""CPU tests of the dosage path's host side: the BGEN container plumbing (dissect_b200/bgenio.py) and the oracle's
restatement of the reference's per-variant standardisation (genotypebgen.cpp:183-282, 723-779).
Parity of this row is UNPINNED: the reference's BGEN reader needs the un-vendored genfile library, so neither the
reference binary nor a golden vector exists for it; the checks below are against the published container layout and
an independent loop-for-loop restatement of the reference's source."""
import math

import numpy as np
import pytest

from dissect_b200 import bgenio
from oracle import grm_oracle as orc


def synth_probs(m, n, seed, denom, missing=0.05):
    rng = np.random.default_rng(seed)
    f = rng.uniform(0.05, 0.5, size=m)
    p_aa = np.empty((m, n), np.uint64)
    p_ab = np.empty((m, n), np.uint64)
    for k in range(m):
        g = rng.binomial(2, f[k], size=n)          # hard call blurred into a probability triple
        w = rng.dirichlet([6.0, 1.0, 1.0], size=n)
        pr = np.zeros((n, 3))
        for j in range(3):
            pr[np.arange(n), (g + j) % 3] = w[:, j]
        a = np.floor(pr[:, 2 - 2] * denom).astype(np.uint64)   # P(AA) <-> 0 copies of B
        b = np.floor(pr[:, 1] * denom).astype(np.uint64)
        p_aa[k], p_ab[k] = a, np.minimum(b, denom - a)
    miss = rng.random((m, n)) < missing
    return p_aa, p_ab, miss


@pytest.mark.parametrize("layout,bits,compressed", [(2, 8, True), (2, 8, False), (2, 16, True), (2, 5, True),
                                                    (2, 32, False), (1, 16, True), (1, 16, False)])
def test_bgen_roundtrip(tmp_path, layout, bits, compressed):
    m, n = 37, 53
    denom = 32768 if layout == 1 else (1 << bits) - 1
    p_aa, p_ab, miss = synth_probs(m, n, 7 + bits, denom)
    path = str(tmp_path / "t.bgen")
    ids = ["f%d i%d" % (i, i) for i in range(n)]
    bgenio.write_bgen(path, p_aa, p_ab, miss, layout=layout, bits=bits, compressed=compressed, sample_ids=ids)
    hdr = bgenio.read_header(path)
    assert (hdr.n_variants, hdr.n_samples, hdr.layout, hdr.compression) == (m, n, layout, int(compressed))
    assert hdr.sample_ids == ids
    blocks = list(bgenio.iter_dosage_blocks(path, block=10))
    assert [len(b[0]) for b in blocks] == [10, 10, 10, 7]
    meta = sum((b[0] for b in blocks), [])
    assert [t[2] for t in meta] == ["rs%d" % k for k in range(m)] and meta[3][1] == 1003 and meta[0][3:] == ("A", "G")
    dos = np.vstack([b[1] for b in blocks])
    mis = np.vstack([b[2] for b in blocks])
    assert np.array_equal(mis != 0, miss)
    a = p_aa.astype(np.float64) / denom
    b = p_ab.astype(np.float64) / denom
    if layout == 1:
        c = (32768 - p_aa - p_ab).astype(np.float64) / denom
    else:
        c = 1.0 - (a + b)
    want = b + 2.0 * c
    assert np.array_equal(dos[~miss], want[~miss])
    assert np.all(dos[miss] == (0.0 if layout == 1 else 2.0))  # stored zeros: layout 2 implies P(BB) = 1 there


def test_bgen_keep_and_extract(tmp_path):
    m, n = 20, 31
    p_aa, p_ab, miss = synth_probs(m, n, 3, 255)
    path = str(tmp_path / "k.bgen")
    bgenio.write_bgen(path, p_aa, p_ab, miss)
    full = list(bgenio.iter_dosage_blocks(path, block=100))[0]
    keep = (np.arange(n) % 3 != 1).astype(np.int32)
    snps = {"rs2", "rs5", "rs19"}
    sub = list(bgenio.iter_dosage_blocks(path, keep_mask=keep, keep_snps=snps, block=100))[0]
    rows = [2, 5, 19]
    assert [t[2] for t in sub[0]] == ["rs2", "rs5", "rs19"]
    assert np.array_equal(sub[1], full[1][rows][:, keep != 0])
    assert np.array_equal(sub[2], full[2][rows][:, keep != 0])


def test_bgen_rejects_what_the_reference_rejects(tmp_path):
    path = str(tmp_path / "bad.bgen")
    p = np.full((1, 4), 100, np.uint64)
    bgenio.write_bgen(path, p, p, compressed=False)
    blob = bytearray(open(path, "rb").read())
    bad = bytes(blob).replace(b"bgen", b"nope")
    open(path, "wb").write(bad)
    with pytest.raises(bgenio.BgenError, match="not a valid BGEN"):
        bgenio.read_header(path)
    # phased flag (the byte before the bit depth, uncompressed layout 2): "can only load unphased genotypes"
    blob[-(8 + 2)] = 1
    open(path, "wb").write(bytes(blob))
    with pytest.raises(bgenio.BgenError, match="unphased"):
        list(bgenio.iter_dosage_blocks(path))


def reference_loops(d, present):
    """ProbabilitiesSetter::finalise, genotypebgen.cpp:239-282, statement by statement for one row (pure Python)."""
    n = len(d)
    mean = 0.0
    for i in range(n):            # set_value: *mean += rowGenotypes[indIdx]  (only reached for present samples in
        if present[i]:            # layout 2; layout-1 missing samples add an exact 0)
            mean += d[i]
    nn = sum(1 for i in range(n) if present[i])
    mean = mean / float(nn)
    g = list(d)
    std = 0.0
    for i in range(n):
        if present[i]:
            g[i] = g[i] - mean
            std += g[i] * g[i]
    std = math.sqrt(std / float(n))
    for i in range(n):
        g[i] = g[i] / std if present[i] else 0.0
    return mean, std, g


def test_oracle_standardise_follows_the_reference_loops():
    rng = np.random.default_rng(11)
    m, n = 25, 41
    d = rng.integers(0, 511, size=(m, n)).astype(np.float64) / 255.0
    present = rng.random((m, n)) > 0.1
    st = orc.dosage_standardise(d, present)
    for k in range(m):
        mean, std, g = reference_loops(d[k], present[k])
        assert st["mean"][k] == mean and st["sd"][k] == std
        assert np.array_equal(st["z"][k], np.array(g))
        assert st["p2"][k] == mean / 2.0 and st["p1"][k] == 1.0 - mean / 2.0


def test_oracle_zero_variance_rows_are_dropped():
    d = np.array([[1.0, 1.0, 1.0, 1.0], [0.0, 1.0, 2.0, 1.0], [0.5, 0.5, 2.0, 0.5]])
    present = np.array([[1, 1, 1, 1], [1, 1, 1, 1], [1, 1, 0, 1]])
    r = orc.dosage_grm(d, present)
    st = r["stats"]
    assert st["sd"][0] == 0.0 and st["sd"][2] == 0.0 and st["sd"][1] > 0
    assert not st["present"][0].any() and not st["present"][2].any()
    low = np.tril(np.ones((4, 4), bool))
    assert np.all(r["N"][low] == 1.0)                       # only SNP 1 is left
    z = st["z"][1]
    assert np.allclose(r["S"][low], np.outer(z, z)[low])
    assert abs(z.mean()) < 1e-15 and abs((z * z).sum() / 4 - 1.0) < 1e-15


# ------------------------------------------------------------------------------- C++ host mirror (no GPU needed)
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SELFTEST = os.path.join(ROOT, "dissect_b200", "host", "selftest")
needs_selftest = pytest.mark.skipif(not os.path.exists(SELFTEST), reason="host selftest not built (build())")


def run_selftest_bgen(args):
    r = subprocess.run([SELFTEST, "bgen"] + args, capture_output=True, text=True)
    inds = [tuple(l.split()[1:]) for l in r.stdout.splitlines() if l.startswith("IND ")]
    snps = [tuple(l.split()[1:]) for l in r.stdout.splitlines() if l.startswith("SNP ")]
    rows = [[(float.fromhex(t.split(":")[0]), int(t.split(":")[1])) for t in l.split()[1:]]
            for l in r.stdout.splitlines() if l.startswith("ROW")]
    return r, inds, snps, rows


@needs_selftest
@pytest.mark.parametrize("layout,bits,compressed,with_ids", [(2, 8, True, True), (2, 16, False, True), (2, 11, True, False),
                                                             (2, 32, True, True), (1, 16, True, True), (1, 16, False, False)])
def test_cpp_bgen_reader_equals_python_reader(tmp_path, layout, bits, compressed, with_ids):
    """The host mirror's Genotype(bgen) + streamBGEN (what Kernel::compute pushes to the engines) delivers bit-identical
    dosage rows, the same missing flags, identifiers and --keep / --extract selection as dissect_b200/bgenio.py."""
    m, n = 23, 29
    denom = 32768 if layout == 1 else (1 << bits) - 1
    p_aa, p_ab, miss = synth_probs(m, n, 40 + bits, denom)
    prefix = str(tmp_path / "g")
    ids = ["s%d" % i for i in range(n)]
    bgenio.write_bgen(prefix + ".bgen", p_aa, p_ab, miss, layout=layout, bits=bits, compressed=compressed,
                      sample_ids=ids if with_ids else None)
    if not with_ids:
        with open(prefix + ".sample", "w") as f:
            f.write("ID_1 ID_2 missing\n0 0 0\n")
            for i in range(n):
                f.write("fam%d ind%d 0\n" % (i, i))
    fam = (lambda i: "s%d" % i) if with_ids else (lambda i: "fam%d" % i)
    ind = (lambda i: "s%d" % i) if with_ids else (lambda i: "ind%d" % i)
    keep_idx = [i for i in range(n) if i % 4 != 1]
    with open(prefix + ".keep", "w") as f:
        for i in keep_idx:
            f.write("%s %s\n" % (fam(i), ind(i)))
    extract = ["rs%d" % k for k in range(m) if k % 3 != 0]
    with open(prefix + ".extract", "w") as f:
        f.write("\n".join(extract) + "\n")
    l1 = ["--bgen-l1"] if layout == 1 else []
    for sel in ([], ["--keep", prefix + ".keep", "--extract", prefix + ".extract"]):
        r, inds, snps, rows = run_selftest_bgen(["--bgen", prefix] + l1 + sel)
        assert r.returncode == 0, r.stdout + r.stderr
        cols = keep_idx if sel else list(range(n))
        keep_mask = np.zeros(n, np.int32)
        keep_mask[cols] = 1
        blocks = list(bgenio.iter_dosage_blocks(prefix + ".bgen", keep_mask=keep_mask, keep_snps=set(extract) if sel else None,
                                                block=1 << 20))
        meta, dos, mis = blocks[0]
        assert inds == [(fam(i), ind(i)) for i in cols]
        assert snps == [(t[0], t[2], str(t[1]), t[3], t[4]) for t in meta]
        got_d = np.array([[v for v, _ in row] for row in rows])
        got_m = np.array([[q for _, q in row] for row in rows])
        assert got_d.shape == dos.shape
        assert np.array_equal(got_d.view(np.uint64), dos.view(np.uint64))
        assert np.array_equal(got_m, mis)


@needs_selftest
def test_cpp_bgen_reader_refusals(tmp_path):
    prefix = str(tmp_path / "g")
    p = np.full((2, 5), 60, np.uint64)
    bgenio.write_bgen(prefix + ".bgen", p, p, layout=1, sample_ids=["a", "b", "c", "d", "e"])
    r, *_ = run_selftest_bgen(["--bgen", prefix])                 # genotypebgen.cpp:337-343: layout 1 needs --bgen-l1
    assert r.returncode != 0 and "--bgen-l1" in r.stdout + r.stderr
    bgenio.write_bgen(prefix + ".bgen", p, p, layout=2)           # no identifiers and no .sample file
    r, *_ = run_selftest_bgen(["--bgen", prefix])
    assert r.returncode != 0
    r, *_ = run_selftest_bgen(["--bgen", str(tmp_path / "nothere")])
    assert r.returncode != 0 and "Unable to read the file" in r.stdout + r.stderr


def fnv1a64(b):
    h = 1469598103934665603
    for x in bytes(b):
        h = ((h ^ x) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@needs_selftest
@pytest.mark.parametrize("layout,bits,compressed", [(2, 8, True), (2, 13, True), (2, 16, False), (1, 16, True), (1, 16, False)])
def test_cpp_block_stream_equals_python_block_stream(tmp_path, layout, bits, compressed):
    """Genotype::streamBGENBlocks (parallel inflate; what the CLI hands to dgrm_push_bgen_blocks_host) delivers the same
    inflated blocks, byte for byte and in file order, as bgenio.iter_raw_blocks, with --extract applied."""
    m, n = 150, 61
    denom = 32768 if layout == 1 else (1 << bits) - 1
    p_aa, p_ab, miss = synth_probs(m, n, 70 + bits, denom)
    prefix = str(tmp_path / "b")
    bgenio.write_bgen(prefix + ".bgen", p_aa, p_ab, miss, layout=layout, bits=bits, compressed=compressed,
                      sample_ids=["s%d" % i for i in range(n)])
    extract = ["rs%d" % k for k in range(m) if k % 5 != 2]
    with open(prefix + ".extract", "w") as f:
        f.write("\n".join(extract) + "\n")
    l1 = ["--bgen-l1"] if layout == 1 else []
    for sel, keep_snps in (([], None), (["--extract", prefix + ".extract"], set(extract))):
        r = subprocess.run([SELFTEST, "bgenblocks", "--bgen", prefix] + l1 + sel, capture_output=True, text=True)
        assert r.returncode == 0, r.stdout + r.stderr
        got = [(int(l.split()[1]), int(l.split()[2], 16)) for l in r.stdout.splitlines() if l.startswith("BLOCK")]
        want = []
        for _, data, off, _ in bgenio.iter_raw_blocks(prefix + ".bgen", keep_snps=keep_snps, block=40):
            want += [(int(off[k + 1] - off[k]), fnv1a64(data[off[k]:off[k + 1]])) for k in range(len(off) - 1)]
        assert got == want


@needs_selftest
def test_cpp_bgen_reader_survives_truncated_and_corrupt_files(tmp_path):
    """Length fields are checked against the file size before they are believed: a cut or corrupted file ends in the
    reader's error, never in a huge allocation or a crash."""
    m, n = 12, 33
    p_aa, p_ab, miss = synth_probs(m, n, 5, 255)
    prefix = str(tmp_path / "t")
    bgenio.write_bgen(prefix + ".bgen", p_aa, p_ab, miss, sample_ids=["s%d" % i for i in range(n)])
    blob = open(prefix + ".bgen", "rb").read()
    cases = [blob[: len(blob) // 2], blob[: len(blob) - 3], blob[:40]]
    first = int.from_bytes(blob[:4], "little") + 4           # first variant block: its id length field
    corrupt = bytearray(blob)
    corrupt[first + 2 + 3 + 2 + 3 + 2 + 1 + 4 + 2: first + 2 + 3 + 2 + 3 + 2 + 1 + 4 + 2 + 4] = (0xFFFFFFF0).to_bytes(4, "little")
    cases.append(bytes(corrupt))                             # allele-1 length field of variant 0 = 4 GB
    for k, data in enumerate(cases):
        open(prefix + ".bgen", "wb").write(data)
        for mode in ("bgen", "bgenblocks"):
            r = subprocess.run([SELFTEST, mode, "--bgen", prefix], capture_output=True, text=True, timeout=60)
            assert r.returncode == 1, (k, mode, r.returncode, r.stderr[-300:])
            assert "Error:" in r.stdout + r.stderr


# ------------------------------------------------------------------------------- property tests of the container
from hypothesis import given, settings, strategies as st_


@settings(max_examples=40, deadline=None)
@given(bits=st_.integers(1, 32), n=st_.integers(1, 70), m=st_.integers(1, 6), compressed=st_.booleans(),
       seed=st_.integers(0, 2 ** 31 - 1))
def test_bgen_layout2_roundtrip_any_bit_depth(tmp_path_factory, bits, n, m, compressed, seed):
    """Every bit depth the format allows (1..32), ragged sample counts: numerators survive write -> read exactly, the
    dosage is the documented expression of them, the raw block stream carries the same bytes the row reader decodes."""
    rng = np.random.default_rng(seed)
    denom = (1 << bits) - 1
    p_aa = rng.integers(0, denom + 1, size=(m, n), dtype=np.uint64)
    p_ab = np.array([[rng.integers(0, denom - int(a) + 1) for a in row] for row in p_aa], dtype=np.uint64)
    miss = rng.random((m, n)) < 0.2
    path = str(tmp_path_factory.mktemp("h") / "p.bgen")
    bgenio.write_bgen(path, p_aa, p_ab, miss, layout=2, bits=bits, compressed=compressed)
    _, dos, mis = list(bgenio.iter_dosage_blocks(path, block=1 << 20))[0]
    assert np.array_equal(mis != 0, miss)
    a, b = p_aa.astype(np.float64) / denom, p_ab.astype(np.float64) / denom
    want = b + 2.0 * (1.0 - (a + b))
    assert np.array_equal(dos[~miss], want[~miss])
    _, data, off, hdr = list(bgenio.iter_raw_blocks(path, block=1 << 20))[0]
    assert hdr.layout == 2 and len(off) == m + 1
    for k in range(m):
        blk = bytes(data[off[k]:off[k + 1]])
        assert blk[9 + n] == bits and int.from_bytes(blk[:4], "little") == n
        ints = bgenio._unpack_bits(blk[10 + n:], 2 * n, bits).reshape(n, 2)
        keep = ~miss[k]
        assert np.array_equal(ints[keep, 0], p_aa[k][keep]) and np.array_equal(ints[keep, 1], p_ab[k][keep])
