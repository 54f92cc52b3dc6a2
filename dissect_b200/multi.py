"""Multi-GPU GRM build: one process per GPU, the lower triangle split into tile sets, the packed 2-bit SNP rows
of every chunk group all-gathered over NVLink (the path's one exchange step; no reduction).

This module IS the multi-GPU path: `bench.py --gpus N`, `tools/multi_gpu_check.py` and the `-m gpu` multi-rank
tests all drive `MultiGpuGrm.build`, so what is verified is what is timed.

Replaces, for G ranks, what the reference does with BLACS sends of fp64 genotype blocks (genotype.cpp:639-707,
16 bytes per genotype per destination) followed by pdsyrk_ over the process grid (kernel.cpp:94-100): here every
rank keeps 1/G of each chunk group's packed rows (1/4 byte per genotype), one all-gather makes the group visible to
every engine, and each engine contracts its own tile set (`dgrm_create` rank/world).

Egress (`reduce_block_cyclic`, `write_grm_dat`): the reference gathers / scatters every block through rank 0
(matrix.cpp:2262-2349, 2168-2219); here each engine exports the entries it owns of a destination rank's
block-cyclic local matrix into a zeroed device panel and the panels are summed over NVLink onto the destination
(every entry has exactly one owner, so the sum is exact), or every engine writes its own runs of the .grm.dat.
"""
import math
import os
import time

import numpy as np

from . import _lib  # noqa: F401  (fails loudly when libdgrm.so is missing)
from .engine import GrmEngine, numroc


class ShardPlan:
    """Pure index arithmetic of the exchange: which file rows every rank keeps per chunk group, and which of the
    G * shard gathered rows of a group are real, kept SNPs (shards are padded to `shard` rows).

    Groups are cut by KEPT rows: every group but the last delivers exactly `xchunk` kept SNPs, i.e. whole engine
    chunks, so a chunk never has to wait for the first rows of the next exchange (with an --extract list that drops 2 %
    of the SNPs, groups cut by file rows left every chunk 2 % short until the next group arrived: 0.5 ms of idle MMA
    stream at the head of each GRM on 8 GPUs, profiles/r2_trace_n8_rank0_tail.txt).

    groups[g] = (first file row, [(row0, row1) per rank], row_keep uint8[G * shard])"""

    def __init__(self, m_file, world, chunk_snps, xg=1, keep=None):
        self.m_file, self.world = int(m_file), int(world)
        self.xchunk = int(chunk_snps) * max(1, int(xg))
        keep = np.ones(m_file, np.uint8) if keep is None else np.ascontiguousarray(keep, dtype=np.uint8)
        assert keep.shape == (m_file,)
        self.keep = keep
        self.m_kept = int(np.count_nonzero(keep))
        # file-row boundaries: after every xchunk-th kept row
        kept_idx = np.flatnonzero(keep)
        cuts = [0]
        for k in range(self.xchunk, len(kept_idx), self.xchunk):
            cuts.append(int(kept_idx[k - 1]) + 1)
        if cuts[-1] != m_file:
            cuts.append(m_file)
        spans = [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1) if cuts[i + 1] > cuts[i]]
        self.shard = max([(c1 - c0 + world - 1) // world for c0, c1 in spans] + [1])
        self.groups = []
        for c0, c1 in spans:
            per = (c1 - c0 + world - 1) // world
            ranges, rk = [], np.zeros(world * self.shard, np.uint8)
            for g in range(world):
                a = min(c0 + g * per, c1)
                b = min(a + per, c1)
                ranges.append((a, b))
                if b > a:
                    rk[g * self.shard: g * self.shard + (b - a)] = keep[a:b] != 0
            self.groups.append((c0, ranges, rk))

    def rows_of(self, rank):
        """file rows this rank holds, over all groups"""
        return sum(b - a for _, ranges, _ in self.groups for (a, b) in [ranges[rank]])


def auto_xg(n, world, chunk_snps, m_file, exchange="nccl", budget_bytes=4 << 30):
    """Engine chunks per exchange.  For large cohorts (>= 256 MB of packed rows per chunk) fewer, larger exchanges
    win with either exchange — measured on one 8-GPU box at C3 (r2): copy-engine exchange 722 ms with one chunk per
    exchange, 664 ms with four; NCCL all-gather 700 ms with four (779 ms with one, r1) — for small ones one chunk per
    exchange pipelines best (28k individuals on 2 GPUs: 2.5 % faster than six)."""
    if world == 1:
        return 1
    pitch = ((n + 3) // 4 + 15) // 16 * 16
    if chunk_snps * pitch < (256 << 20):
        return 1
    per_chunk = 2 * (1 + 1.0 / world) * chunk_snps * pitch
    nchunks = max(1, math.ceil(m_file / chunk_snps))
    return int(max(1, min(4, nchunks, budget_bytes // max(1, int(per_chunk)))))


class DeviceRows:
    """Row source: the packed cohort as a CUDA uint8 tensor [m_file, >= ceil(n/4)] on this rank's GPU (a rank
    reads only its own shards of it)."""

    def __init__(self, pk):
        self.pk = pk

    def fill(self, send_buf, a, b, bpr, stream):
        send_buf[: b - a, :bpr].copy_(self.pk[a:b, :bpr])


class HostRows:
    """Row source: packed rows in (pinned) host memory at `ptr`, `pitch` bytes apart — the e2e path: each rank
    uploads only its own shards."""

    def __init__(self, ptr, pitch):
        from cuda.bindings import runtime as cudart
        self.cudart, self.ptr, self.pitch = cudart, int(ptr), int(pitch)

    def fill(self, send_buf, a, b, bpr, stream):
        err, = self.cudart.cudaMemcpy2DAsync(send_buf.data_ptr(), send_buf.stride(0), self.ptr + a * self.pitch,
                                             self.pitch, bpr, b - a,
                                             self.cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream)
        if int(err) != 0:
            raise RuntimeError(f"cudaMemcpy2DAsync failed: {err}")


class MultiGpuGrm:
    """One per rank.  `engine` is this rank's GrmEngine (its tile set); `build` runs the whole GRM."""

    def __init__(self, n, *, rank, world, device, chunk_snps=16384, xg=0, exchange="auto", group=None,
                 engine=None, **engine_kw):
        """exchange: "p2p" = copy-engine peer writes over NVLink (dissect_b200/exchange.py, SM-free, runs underneath the
        contraction), "nccl" = all_gather_into_tensor, "auto" = p2p when every rank can map its peers' buffers
        (torch symmetric memory), else nccl."""
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.n, self.rank, self.world, self.device = int(n), int(rank), int(world), device
        self.chunk, self.xg_arg, self.group = int(chunk_snps), int(xg), group
        self.bpr = (self.n + 3) // 4
        self.pitch = (self.bpr + 15) // 16 * 16     # 16-byte aligned rows: the engine's row gather uses 128-bit copies
        self.eng = engine if engine is not None else GrmEngine(device=device, rank=rank, world=world,
                                                               chunk_snps=chunk_snps, **engine_kw)
        self.eng.set_individuals(self.n)
        # device = "cpu" (with a stand-in engine and a gloo group) runs the same loop in the CPU tests of the host logic
        self._cuda = device != "cpu"
        self._dev = torch.device("cuda", device) if self._cuda else torch.device("cpu")
        if exchange == "auto":
            exchange = "p2p" if (self._cuda and world > 1) else "nccl"
            self._exchange_fallback = True
        else:
            self._exchange_fallback = False
        self.exchange = exchange if world > 1 else "none"
        self.plan = None
        self._bufs = None
        self._xch = None

    # -- planning / buffers -----------------------------------------------------------------------------------
    def prepare(self, m_file, keep=None):
        """Builds the shard plan and (re)allocates the exchange buffers; idempotent for the same cohort."""
        if self.world > 1 and self.exchange == "p2p" and self._xch is None and self._exchange_fallback and not getattr(self, "_p2p_probed", False):
            # decide once, collectively, whether the peer mapping works on this box
            from .exchange import PeerExchange
            self._p2p_probed = True
            self._host_group = self.dist.new_group(backend="gloo")
            probe = PeerExchange(1, 16, self.rank, self.world, self._host_group)
            ok = probe.ok
            probe.close()
            if not ok:
                self.exchange = "nccl"
        xg = self.xg_arg if self.xg_arg > 0 else auto_xg(self.n, self.world, self.chunk, m_file, self.exchange)
        if self.plan is not None and self.plan.m_file == m_file and self.plan.xchunk == self.chunk * xg and (
                (keep is None and self.plan.m_kept == m_file) or (keep is not None and np.array_equal(self.plan.keep, keep))):
            return self.plan
        self.plan = ShardPlan(m_file, self.world, self.chunk, xg, keep)
        self.xg = xg
        if self.world > 1:
            torch = self.torch
            dev = self._dev
            sh = self.plan.shard
            self._close_exchange()
            if self.exchange == "p2p":
                from .exchange import PeerExchange
                if getattr(self, "_host_group", None) is None:
                    self._host_group = self.dist.new_group(backend="gloo")
                self._xch = PeerExchange(sh, self.pitch, self.rank, self.world, self._host_group)
                if not self._xch.ok:
                    raise RuntimeError(f"peer exchange unavailable: {self._xch.error}")
                self._bufs = {"send": [torch.zeros((sh, self.pitch), dtype=torch.uint8, device=dev) for _ in range(3)]}
            else:
                self._bufs = {
                    "send": [torch.zeros((sh, self.pitch), dtype=torch.uint8, device=dev) for _ in range(2)],
                    "gather": [torch.zeros((self.world * sh, self.pitch), dtype=torch.uint8, device=dev) for _ in range(2)],
                }
        return self.plan

    def _close_exchange(self):
        if self._xch is not None:
            self._xch.close()
            self._xch = None

    def close(self):
        self._close_exchange()
        self._bufs = None
        if self.eng is not None:
            self.eng.close()
            self.eng = None

    # -- the GRM build ----------------------------------------------------------------------------------------
    def build(self, source, m_file, keep=None, finish=True):
        """begin -> for every chunk group: this rank's shard -> exchange -> engine -> finish.
        Everything is stream-ordered on torch's current stream; the host never waits inside the loop."""
        plan = self.prepare(m_file, keep)
        eng, torch = self.eng, self.torch
        trace = os.environ.get("DGRM_BUILD_TRACE") and self.rank == 0
        t0 = time.perf_counter()
        eng.begin(max(plan.m_kept, 1))
        stream = torch.cuda.current_stream().cuda_stream if self._cuda else None
        if self.world == 1:
            if isinstance(source, DeviceRows):
                eng.push_bed_device(source.pk.data_ptr(), source.pk.stride(0), m_file, row_keep=keep, stream=stream)
            else:
                eng.push_bed_host(source.ptr, row_keep=keep, pitch=source.pitch, nrows=m_file)
        elif self._xch is not None:
            self._build_p2p(source, plan)
        else:
            sends, gathers = self._bufs["send"], self._bufs["gather"]
            for gi, (c0, ranges, rk) in enumerate(plan.groups):
                # two buffer pairs: group gi+2 reuses pair gi&1.  Its shard copy and all-gather are ordered on
                # `stream` after dgrm_release_inputs of group gi (below), i.e. after the engine's row gather has
                # read the buffer; the engine orders its read after the all-gather through the producer stream.
                send_buf, gather_buf = sends[gi & 1], gathers[gi & 1]
                a, b = ranges[self.rank]
                if b > a:
                    source.fill(send_buf, a, b, self.bpr, stream)
                self.dist.all_gather_into_tensor(gather_buf, send_buf, group=self.group)
                eng.push_bed_device(gather_buf.data_ptr(), gather_buf.stride(0), self.world * plan.shard, row_keep=rk,
                                    stream=stream)
                eng.release_inputs(stream)
        t1 = time.perf_counter()
        if finish:
            eng.finish()
        if trace:
            import sys
            print(f"[multi r0] push phase {1e3 * (t1 - t0):.2f} ms (host time to enqueue), finish {1e3 * (time.perf_counter() - t1):.2f} ms, "
                  f"groups {len(plan.groups)}, exchange {self.exchange}", file=sys.stderr)

    def _build_p2p(self, source, plan):
        """Copy-engine exchange (dissect_b200/exchange.py): shards are posted up to NBUF groups ahead."""
        from .exchange import NBUF
        xch, sends = self._xch, self._bufs["send"]
        cs = xch.copy_stream.cuda_stream
        ng, posted = len(plan.groups), 0

        def post_until(limit):
            nonlocal posted
            while posted < min(ng, limit):
                a, b = plan.groups[posted][1][self.rank]
                sb = sends[posted % len(sends)]
                if isinstance(source, DeviceRows) and source.pk.stride(0) == self.pitch:
                    ptr = source.pk.data_ptr() + a * self.pitch          # contiguous rows at the exchange pitch: no staging
                else:
                    if b > a:
                        with self.torch.cuda.stream(xch.copy_stream):
                            source.fill(sb, a, b, self.bpr, cs)
                    ptr = sb.data_ptr()
                xch.post(ptr, b - a)
                posted += 1

        for gi in range(ng):
            post_until(gi + 1)                 # this group's shard is on its way ...
            xch.consume(self.eng, plan.groups[gi][2])   # ... its decode is queued behind the arrival flags ...
            post_until(gi + NBUF)              # ... and only then does the host run ahead with the next groups

    # -- egress -----------------------------------------------------------------------------------------------
    def gather_tiles(self, which, dst=0):
        """Host-side gather of every rank's (tile coordinates, tiles) onto rank `dst` (small cases / checks)."""
        payload = (self.eng.tile_coords(), self.eng.fetch_tiles(which))
        if self.world == 1:
            return [payload]
        out = [None] * self.world if self.rank == dst else None
        self.dist.gather_object(payload, out, dst=dst, group=self.group)
        return out

    def reduce_block_cyclic(self, which, bs, grid, out=None, uplo="L", panel_cols=0):
        """The reference's block-cyclic Matrix for an R x C process grid with R * C == world, rank r at grid
        position (r // C, r % C) (row-major, communicator.cpp:102, matrix.cpp:719): returns this rank's local
        block (Fortran order, numroc rows x numroc columns; `out` = a caller array / PinnedArray.array to fill).

        For every destination rank d, column panel by column panel: each engine exports the entries it owns of
        d's local block into a zeroed device panel (dgrm_fetch_block_cyclic with a device destination) and the
        panels are summed onto d over NVLink (ncclReduce).  Each entry has exactly one owner: the sum is exact."""
        torch, dist = self.torch, self.dist
        R, C = grid
        assert R * C == self.world, "the process grid must have one rank per engine"
        n = self.eng._n_out if getattr(self.eng, "_n_out", None) else self.n
        my = (self.rank // C, self.rank % C)
        nr_me, nc_me = numroc(n, bs[0], my[0], R), numroc(n, bs[1], my[1], C)
        if out is None:
            out = np.full((nr_me, nc_me), np.nan, dtype=np.float64, order="F")
        assert out.shape == (nr_me, nc_me) and out.dtype == np.float64 and (out.flags.f_contiguous or out.size == 0)
        dev = self._dev
        for d in range(self.world):
            pr, pc = d // C, d % C
            nr, nc = numroc(n, bs[0], pr, R), numroc(n, bs[1], pc, C)
            if nr == 0 or nc == 0:
                continue
            pcw = panel_cols if panel_cols > 0 else max(1, min(nc, (256 << 20) // (8 * nr)))
            for lc0 in range(0, nc, pcw):
                lc1 = min(lc0 + pcw, nc)
                panel = torch.zeros((lc1 - lc0, nr), dtype=torch.float64, device=dev)    # column-major local panel
                if self._cuda:
                    torch.cuda.current_stream().synchronize()    # the engine writes on its own stream
                self.eng.fetch_into(which, panel.data_ptr(), nr, bs=bs, grid=grid, coords=(pr, pc), uplo=uplo,
                                    cols=(lc0, lc1))
                if self.world > 1:
                    dist.reduce(panel, dst=d, op=dist.ReduceOp.SUM, group=self.group)
                if d == self.rank:
                    blk = panel.cpu().numpy()                       # [cols][rows]
                    tgt = out[:, lc0:lc1]
                    if uplo == "B":
                        tgt[...] = blk.T
                    else:
                        # only the requested triangle is defined (pdsyrk_ leaves the rest untouched, matrix.cpp:2679)
                        gr = _local_to_global(np.arange(nr), bs[0], pr, R)[:, None]
                        gc = _local_to_global(np.arange(lc0, lc1), bs[1], pc, C)[None, :]
                        mask = gc <= gr if uplo == "L" else gc >= gr
                        np.copyto(tgt, blk.T, where=mask)
        return out

    def write_grm_dat(self, path):
        """<path> = the reference's .grm.dat (kernel.cpp:944-974), written by all engines in parallel: rank 0
        creates and sizes the file, then every engine writes the runs it owns (dgrm_write_grm_dat_part)."""
        if self.rank == 0:
            self.eng.write_grm_dat_part(path, create=True)
        if self.world > 1:
            self.dist.barrier(group=self.group)
            if self.rank != 0:
                self.eng.write_grm_dat_part(path, create=False)
            self.dist.barrier(group=self.group)


def _local_to_global(l, bs, me, P):
    return ((l // bs) * P + me) * bs + l % bs
