"""CPU: the product's multi-GPU driver (dissect_b200/multi.py) with world_size = 2 and 3 over gloo.  The CUDA engine
is replaced by a stand-in that records the rows it is handed, so the exchange loop that bench.py times on NCCL —
shard plan, double-buffered all-gather, row_keep tables — runs here unchanged and must deliver exactly the kept SNP
rows, in file order, to every rank."""
import ctypes
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dissect_b200 import bedio
from dissect_b200.multi import DeviceRows, MultiGpuGrm, ShardPlan, auto_xg


class RecordingEngine:
    """What MultiGpuGrm needs from GrmEngine; keeps the rows the real engine's row gather would read."""

    def __init__(self):
        self.rows, self.begun, self.finished, self.released = [], 0, 0, 0

    def set_individuals(self, n):
        self.n = n

    def begin(self, m_cap):
        self.rows, self.m_cap = [], m_cap
        self.begun += 1

    def push_bed_device(self, ptr, pitch, nrows, row_keep=None, stream=None):
        buf = np.ctypeslib.as_array((ctypes.c_uint8 * (pitch * nrows)).from_address(ptr)).reshape(nrows, pitch)
        bpr = (self.n + 3) // 4
        for r in range(nrows):
            if row_keep is None or row_keep[r]:
                self.rows.append(buf[r, :bpr].copy())       # copy now: the buffer is reused two groups later

    def release_inputs(self, stream=None):
        self.released += 1

    def finish(self):
        self.finished += 1

    def close(self):
        pass


@pytest.mark.parametrize("m,world,chunk,xg", [(1000, 2, 256, 1), (8192, 8, 8192, 1), (100000, 4, 16384, 3), (5, 2, 256, 1),
                                              (1025, 3, 256, 2)])
def test_shard_plan_covers_every_row_once(m, world, chunk, xg):
    rng = np.random.default_rng(m)
    keep = (rng.random(m) < 0.8).astype(np.uint8)
    plan = ShardPlan(m, world, chunk, xg, keep)
    rows = [r for _, ranges, _ in plan.groups for (a, b) in ranges for r in range(a, b)]
    assert rows == list(range(m))
    assert sum(plan.rows_of(r) for r in range(world)) == m
    kept = []
    for c0, ranges, rk in plan.groups:
        assert rk.shape == (world * plan.shard,)
        for g, (a, b) in enumerate(ranges):
            assert b - a <= plan.shard
            seg = rk[g * plan.shard: (g + 1) * plan.shard]
            assert not seg[b - a:].any()                     # padding rows of a shard are never kept
            kept += [a + i for i in np.nonzero(seg[: b - a])[0]]
    assert kept == list(np.nonzero(keep)[0])
    assert plan.m_kept == int(keep.sum())


def test_auto_exchange_granularity():
    assert auto_xg(20000, 1, 16384, 100000) == 1
    assert auto_xg(28284, 2, 16384, 100000, "nccl") == 1            # small all-gathers: one chunk per exchange
    assert auto_xg(100000, 8, 16384, 500000, "nccl") == 4           # 410 MB per chunk: fewer, larger all-gathers
    assert auto_xg(100000, 8, 16384, 500000, "p2p") == 4            # measured at C3: 664 vs 722 ms
    assert auto_xg(56568, 8, 16384, 100000, "p2p") == 1
    assert auto_xg(100000, 8, 16384, 500000, "nccl", budget_bytes=1 << 20) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, m, chunk, xg, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pk = bedio.synth_packed(n, m, 5, missing=0.02)          # same seed on every rank
    keep = (np.arange(m) % 5 != 2).astype(np.uint8)
    eng = RecordingEngine()
    mg = MultiGpuGrm(n, rank=rank, world=world, device="cpu", chunk_snps=chunk, xg=xg, engine=eng)
    src = DeviceRows(torch.from_numpy(pk))
    ok = True
    for rep in range(3):                                    # buffers are reused across GRMs
        mg.build(src, m, keep)
        got = np.stack(eng.rows) if eng.rows else np.zeros((0, pk.shape[1]), np.uint8)
        ok = ok and np.array_equal(got, pk[keep != 0])
    ok = ok and eng.begun == 3 and eng.finished == 3 and eng.released == 3 * len(mg.plan.groups)
    np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(ok), len(mg.plan.groups)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,m,chunk,xg", [(2, 700, 256, 1), (3, 1500, 256, 2), (2, 100, 256, 0)])
def test_multi_driver_gloo(world, m, chunk, xg, tmp_path):
    n = 301                                                  # ceil(n/4) = 76 bytes: not a multiple of the 16-byte pitch
    mp.spawn(_worker, args=(world, _free_port(), n, m, chunk, xg, str(tmp_path)), nprocs=world, join=True)
    res = [np.load(str(tmp_path / f"ok{r}.npy")) for r in range(world)]
    assert all(r[0] == 1 for r in res)                       # every rank was handed exactly the kept rows, in order
