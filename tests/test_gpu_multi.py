"""GPU: multi-engine behaviour through the C ABI — partial owners (rank r of a world), the multi-engine egress
(block-cyclic panels in device memory, the parallel .grm.dat writer), the fp32 transport of streamed A values —
and, when the box has two or more GPUs, the real 2-rank run of tools/multi_gpu_check.py (NCCL)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from dissect_b200 import GrmEngine, bedio
from dissect_b200.engine import PinnedArray, numroc
from oracle import grm_oracle as orc
from tests.helpers import lower_mask

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build(pk, n, **kw):
    eng = GrmEngine(**kw)
    eng.set_individuals(n)
    eng.begin(pk.shape[0])
    eng.push_bed_host(np.ascontiguousarray(pk))
    eng.finish()
    return eng


@pytest.mark.parametrize("pinned", [True, False])
@pytest.mark.parametrize("rank,world", [(1, 2), (0, 3), (2, 3)])
def test_partial_owner_bound_output_writes_only_owned_entries(pinned, rank, world):
    """an engine that does not own the first raster row group (rank > 0, n_pad > 2048) with a block-cyclic
    binding: dgrm_finish must fill exactly the entries of its own tiles and leave everything else alone."""
    n, m = 2300, 600
    pk = bedio.synth_packed(n, m, 11, missing=0.02)
    whole = build(pk, n, chunk_snps=256)
    ref = {k: whole.fetch_block_cyclic(k, bs=(n, n)) for k in ("A", "N")}
    whole.close()
    eng = GrmEngine(chunk_snps=256, rank=rank, world=world)
    eng.set_individuals(n)
    owned = np.zeros((n, n), bool)
    for rb, cp in eng.tile_coords():
        owned[rb * 128: rb * 128 + 128, cp * 256: cp * 256 + 256] = True
    owned &= lower_mask(n)
    assert owned.any() and not owned[:2048, :2048].all()
    bufs = {}
    for which in ("A", "N"):
        if pinned:
            pa = PinnedArray((n, n), np.float64, order="F")
            pa.array[:] = -7.0
            bufs[which] = (pa, pa.array, pa.ptr)
        else:
            arr = np.full((n, n), -7.0, order="F")
            bufs[which] = (arr, arr, arr.ctypes.data)
        eng.bind_output(which, bufs[which][2], n, bs=(n, n), uplo="L")
    eng.begin(m)
    eng.push_bed_host(pk)
    eng.finish()
    for which in ("A", "N"):
        got = bufs[which][1]
        assert np.array_equal(got[owned], ref[which][owned]), which
        assert np.all(got[~owned] == -7.0), which
    # the fetch afterwards obeys the same rule, on a 2-D grid as well
    for bs, grid, coords in (((64, 48), (2, 2), (1, 0)), ((n, n), (1, 1), (0, 0))):
        rows = [g for g in range(n) if (g // bs[0]) % grid[0] == coords[0]]
        cols = [g for g in range(n) if (g // bs[1]) % grid[1] == coords[1]]
        got = eng.fetch_block_cyclic("A", bs=bs, grid=grid, coords=coords, uplo="L", fill=-7.0)
        sub_owned = owned[np.ix_(rows, cols)]
        assert np.array_equal(got[sub_owned], ref["A"][np.ix_(rows, cols)][sub_owned])
        assert np.all(got[~sub_owned] == -7.0)
    eng.close()


def test_engines_of_a_partition_fill_one_buffer_in_turn_and_write_one_file(tmp_path):
    """world = 3 engines emulated on one GPU: each fetches into the same caller buffer (owned entries only) and writes
    its own runs of one .grm.dat; both equal the single-engine result."""
    n, m = 1300, 500
    pk = bedio.synth_packed(n, m, 23, missing=0.03)
    whole = build(pk, n, chunk_snps=256)
    ref = {k: whole.fetch_block_cyclic(k, bs=(n, n)) for k in ("A", "N")}
    ref_path = str(tmp_path / "one.grm.dat")
    whole.write_grm_dat(ref_path)
    whole.close()
    path = str(tmp_path / "parts.grm.dat")
    buf = {k: np.full((n, n), np.nan, order="F") for k in ("A", "N")}
    for r in range(3):
        e = build(pk, n, chunk_snps=256, rank=r, world=3)
        with pytest.raises(Exception):
            e.write_grm_dat(path)                       # the whole-triangle writer refuses a partial owner
        e.write_grm_dat_part(path, create=(r == 0))
        for k in buf:
            e.fetch_into(k, buf[k].ctypes.data, n, bs=(n, n))
        e.close()
    low = lower_mask(n)
    for k in buf:
        assert np.array_equal(buf[k][low], ref[k][low]) and np.isnan(buf[k][~low]).all()
    assert open(path, "rb").read() == open(ref_path, "rb").read()


def test_block_cyclic_panels_in_device_memory():
    """dgrm_fetch_block_cyclic_cols into DEVICE memory (the multi-GPU egress panel): same values as the host fetch,
    column range respected, nothing else written."""
    import torch
    n, m = 900, 400
    pk = bedio.synth_packed(n, m, 31, missing=0.02)
    eng = build(pk, n, chunk_snps=256, rank=1, world=2)
    bs, grid, coords = (50, 70), (2, 2), (1, 1)
    nr, nc = numroc(n, bs[0], coords[0], grid[0]), numroc(n, bs[1], coords[1], grid[1])
    want = eng.fetch_block_cyclic("A", bs=bs, grid=grid, coords=coords, uplo="B", fill=0.0)
    lc0, lc1 = 37, 181
    panel = torch.zeros((lc1 - lc0, nr), dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    eng.fetch_into("A", panel.data_ptr(), nr, bs=bs, grid=grid, coords=coords, uplo="B", cols=(lc0, lc1))
    assert np.array_equal(panel.cpu().numpy().T, want[:, lc0:lc1])
    assert np.any(want[:, lc0:lc1] != 0)
    eng.close()


@pytest.mark.parametrize("tiles", [False, True])
def test_fp32_transport_of_streamed_A(tiles):
    """cfg.a_transport = fp32: streamed A arrives as doubles within 2^-24 relative of the fp64 value, the diagonal
    exactly, N untouched; fewer bytes cross PCIe."""
    n, m = 1000, 700
    pk = bedio.synth_packed(n, m, 41, missing=0.02)
    moved = {}
    res = {}
    for tr in ("fp64", "fp32"):
        eng = GrmEngine(chunk_snps=256, a_transport=tr, rank=1 if tiles else 0, world=2 if tiles else 1)
        eng.set_individuals(n)
        if tiles:
            A = PinnedArray((eng.num_tiles, 256, 128), np.float64)
            N = PinnedArray((eng.num_tiles, 256, 128), np.float64)
            eng.bind_tiles("A", A.ptr)
            eng.bind_tiles("N", N.ptr)
        else:
            A = PinnedArray((n, n), np.float64, order="F")
            N = PinnedArray((n, n), np.float64, order="F")
            eng.bind_output("A", A.ptr, n, bs=(n, n))
            eng.bind_output("N", N.ptr, n, bs=(n, n))
        eng.begin(m)
        eng.push_bed_host(pk)
        eng.finish()
        moved[tr] = eng.last_d2h_bytes
        if tiles:
            res[tr] = (A.array.copy(), N.array.copy(), eng.fetch_tiles("A"), eng.tile_coords())
        else:
            res[tr] = (A.array.copy(), N.array.copy(), eng.fetch_block_cyclic("A", bs=(n, n)), None)
        eng.close()
    A64, N64, F64, rc = res["fp64"]
    A32, N32, F32, _ = res["fp32"]
    assert np.array_equal(N64, N32)
    if tiles:
        mask = np.ones(A64.shape, bool)
        assert np.array_equal(A64, F64)
        diag = np.zeros(A64.shape, bool)
        for t, (rb, cp) in enumerate(rc):
            for i in range(max(rb * 128, cp * 256), min(rb * 128 + 128, cp * 256 + 256, n)):
                diag[t, i - cp * 256, i - rb * 128] = True
    else:
        mask = lower_mask(n)
        diag = np.eye(n, dtype=bool)
        assert np.array_equal(A64[mask], F64[mask])
    rel = np.abs(A32[mask] - A64[mask]) / np.maximum(np.abs(A64[mask]), 1e-300)
    assert np.nanmax(rel) <= 2.0 ** -24
    assert np.any(A32[mask] != A64[mask])                   # it really went through fp32
    assert diag.any() and np.array_equal(A32[diag], A64[diag])   # exact fp64 diagonal
    assert 0 < moved["fp32"] < moved["fp64"]


def test_store_bed_device_equals_host_store():
    import torch
    n, m = 333, 500
    pk = bedio.synth_packed(n, m, 51, missing=0.02)
    res = []
    for dev in (False, True):
        eng = GrmEngine(chunk_snps=256)
        eng.set_individuals(n)
        if dev:
            t = torch.from_numpy(pk).cuda()
            eng.store_bed_device(t.data_ptr(), t.stride(0), m, stream=torch.cuda.current_stream().cuda_stream)
        else:
            eng.store_bed_host(pk)
        assert eng.stored_snps == m
        eng.begin(m)
        eng.push_stored(0, m)
        eng.finish()
        res.append((eng.fetch_block_cyclic("S", bs=(n, n)), eng.fetch_block_cyclic("N", bs=(n, n))))
        eng.close()
    low = lower_mask(n)
    assert np.array_equal(res[0][0][low], res[1][0][low]) and np.array_equal(res[0][1][low], res[1][1][low])


def _gpu_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("exchange", ["nccl"])
def test_two_rank_run_matches_oracle(exchange, tmp_path):
    """The real thing: 2 ranks, NCCL, the product's exchange loop (dissect_b200/multi.py), every check of
    tools/multi_gpu_check.py.  Needs two GPUs; the 1-GPU boxes skip it (the emulated-world tests above still run)."""
    if _gpu_count() < 2:
        pytest.skip("needs two GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tools", "multi_gpu_check.py"), "--exchange", exchange, "--out",
           str(tmp_path)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0 and "ALL PASS" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("tiles", [False, True])
def test_pair_count_transport_fp64_equals_narrow(tiles):
    """cfg.n_transport: doubles straight from the copy engine vs 16-bit deficiencies widened by host threads — the
    caller sees the same exact counts either way; the direct form moves 4x the bytes and needs no host work."""
    n, m = 900, 600
    pk = bedio.synth_packed(n, m, 43, missing=0.03)
    res, moved = {}, {}
    for tr in ("narrow", "fp64"):
        eng = GrmEngine(chunk_snps=256, n_transport=tr, rank=0, world=2 if tiles else 1)
        eng.set_individuals(n)
        if tiles:
            N = PinnedArray((eng.num_tiles, 256, 128), np.float64)
            eng.bind_tiles("N", N.ptr)
        else:
            N = PinnedArray((n, n), np.float64, order="F")
            N.array[:] = -1.0
            eng.bind_output("N", N.ptr, n, bs=(n, n))
        eng.begin(m)
        eng.push_bed_host(pk)
        eng.finish()
        moved[tr] = eng.last_d2h_bytes
        want = eng.fetch_tiles("N") if tiles else eng.fetch_block_cyclic("N", bs=(n, n))
        got = N.array.copy()
        if tiles:
            assert np.array_equal(got, want)
        else:
            low = lower_mask(n)
            assert np.array_equal(got[low], want[low])
        res[tr] = got
        eng.close()
    if tiles:
        assert np.array_equal(res["narrow"], res["fp64"])
    else:
        low = lower_mask(n)
        assert np.array_equal(res["narrow"][low], res["fp64"][low])
    assert moved["fp64"] > 2 * moved["narrow"] > 0


def test_eight_engine_compact_partition_covers_triangle_once():
    """world = 8 emulated on one GPU: the macro-block raster (4 x 4 blocks, diagonal halves paired) gives every
    engine a compact region, bed_expand writes operands only for the individuals that region touches, and the union
    of the eight tile sets is the single-engine GRM, value for value."""
    from dissect_b200.engine import assemble_from_tiles
    from dissect_b200.partition import tiles_of_rank
    n, m = 2300, 300                      # T = 9 supertiles per side, macro blocks of 3
    pk = bedio.synth_packed(n, m, 61, missing=0.02)
    whole = build(pk, n, chunk_snps=256)
    ref = {k: whole.fetch_block_cyclic(k, bs=(n, n)) for k in ("A", "N")}
    whole.close()
    sets = {"A": [], "N": []}
    seen = set()
    for r in range(8):
        e = build(pk, n, chunk_snps=256, rank=r, world=8)
        rc = e.tile_coords()
        assert sorted(map(tuple, rc.tolist())) == sorted(tiles_of_rank(n, r, 8))     # host mirror == engine
        for t in map(tuple, rc.tolist()):
            assert t not in seen
            seen.add(t)
        for k in sets:
            sets[k].append((rc, e.fetch_tiles(k)))
        e.close()
    low = lower_mask(n)
    for k in sets:
        merged = assemble_from_tiles(n, sets[k])
        assert np.array_equal(merged[low], ref[k][low]), k
