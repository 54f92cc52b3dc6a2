#!/usr/bin/env python
"""bench.py — GRM build throughput (pair·SNP/s) of the B200-native `--make-grm` hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU implementation

A "step" is one complete GRM build (decode + per-SNP stats + standardise + Z'Z + pair counts) over the
synthetic cohort named in `config.workload`.  `value` times K steps with the packed .bed rows already
resident in HBM; `e2e` times the same build through the C ABI with HOST buffers: pinned .bed rows in,
block-cyclic A and N (lower triangles) out into pinned Matrix::m-style arrays, every step.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

# ---- workload family -----------------------------------------------------------------------------------
# N = 1: BASELINE.json configs[1]  (20,000 individuals x 100,000 SNPs, 1% missing, MAF filter = extract list)
# N > 1: same SNP set, individuals scaled by sqrt(N) so that the pair work per GPU stays fixed (weak scaling);
#        N = 8 gives 56,568 individuals, between configs[1] and configs[2].
BASE_N, BASE_M, MISSING, MAF_LO, MAF_HI, MAF_MIN, SEED = 20000, 100000, 0.01, 0.001, 0.5, 0.01, 2
# bounded CPU sample of the same workload for the reference arm (about 10-30 s of host work)
CPU_N, CPU_M = 4000, 16000


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d["bf16_tflops_sustained"],
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler(threading.Thread):
    """Samples SM clock, power and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.sm, self.power, self.reasons = [], [], 0
        self.max_sm = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            # CUDA_VISIBLE_DEVICES may remap ordinals; the boxes expose GPUs in order, so index == ordinal
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and vis.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        while not self._stop_evt.is_set():
            try:
                self.sm.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                self.reasons |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        if self.is_alive():
            self.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_sm, "reasons": [], "samples": len(self.sm)}
        if self.sm:
            s = sorted(self.sm)
            out["sm_mhz"] = s[len(s) // 2]
            out["sm_mhz_min"] = s[0]
            out["power_w_max"] = max(self.power) if self.power else None
        if self.ok:
            nv = self.nv
            names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
            out["reasons"] = [k for k, bit in names.items() if self.reasons & bit]
        return out


def reference_step(workdir, threads):
    """One step of the reference arm: the reference's own Genotype load + Kernel constructor
    (oracle/_ref/grmtime, built from the unmodified sources) on the bounded CPU sample."""
    from oracle import ref_runner
    exe = os.path.join(ROOT, "oracle", "_ref", "grmtime")
    cmd = [exe, "--make-grm", "--bfile", os.path.join(workdir, "cpu"), "--out", os.path.join(workdir, "cpu_out"),
           "--extract", os.path.join(workdir, "cpu.extract")]
    env = dict(os.environ, OMP_NUM_THREADS=str(threads), OPENBLAS_NUM_THREADS=str(threads))
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=3600)
    wall = time.time() - t0
    mt = re.search(r"GRMTIME n (\d+) m (\d+) threads (\d+) load_s ([\d.]+) compute_s ([\d.]+)", r.stdout)
    if r.returncode != 0 or not mt:
        raise RuntimeError("reference run failed: " + r.stdout[-2000:] + r.stderr[-2000:])
    n, m = int(mt.group(1)), int(mt.group(2))
    load_s, comp_s = float(mt.group(4)), float(mt.group(5))
    del ref_runner
    return {"n": n, "m": m, "load_s": load_s, "compute_s": comp_s, "wall_s": wall, "W": n * (n + 1) / 2 * m}


def make_cpu_sample(workdir, m_file=CPU_M):
    """Writes the bounded CPU sample (same distributions / filter as the GPU workload) as .bed/.bim/.fam."""
    from dissect_b200 import bedio
    pk, maf = bedio.synth_packed(CPU_N, m_file, SEED, maf_lo=MAF_LO, maf_hi=MAF_HI, missing=MISSING, return_maf=True)
    bedio.write_bed(os.path.join(workdir, "cpu"), pk, CPU_N)
    with open(os.path.join(workdir, "cpu.extract"), "w") as f:
        for k in np.nonzero(maf >= MAF_MIN)[0]:
            f.write(f"rs{k + 1}\n")


def cpu_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def sample_desc(r, threads):
    return (f"reference Genotype load + Kernel ctor (oracle/_ref/grmtime, unmodified reference sources, 1 MPI rank "
            f"on a 1x1 BLACS shim, OpenBLAS dsyrk, {threads} threads) on n={r['n']} x m={r['m']} kept SNPs of the same "
            f"synthetic family (1% missing, MAF>=0.01 extract list): load {r['load_s']:.2f}s + compute "
            f"{r['compute_s']:.2f}s; the full config needs >= 42 GB and ~10 min on one rank")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    exe = os.path.join(ROOT, "oracle", "_ref", "grmtime")
    threads = cpu_threads()
    base = {"impl": "reference", "metric": "grm_pair_snp_per_s", "unit": "pair*SNP/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic"}
    # the sample shrinks with K so that the whole --steps K --warmup W run stays within a few minutes
    runs = args.steps + min(args.warmup, 1)
    m_s = max(2000, min(CPU_M, int(CPU_M * 6 / max(runs, 1))))
    if not os.path.exists(exe):
        # the reference-built oracle travels prebuilt; without it fall back to the plain-C port
        from oracle import c_oracle
        from dissect_b200 import bedio
        if not c_oracle.available():
            base.update({"unavailable": "neither oracle/_ref/grmtime nor oracle/_ref/libgrm_oracle.so is built"})
            emit(base)
            return
        pk = bedio.synth_packed(CPU_N // 2, m_s // 4, SEED, maf_lo=0.05, maf_hi=MAF_HI, missing=MISSING)
        res = []
        for i in range(runs):
            t0 = time.time()
            c_oracle.grm(pk, CPU_N // 2)
            if i >= runs - args.steps:
                res.append({"n": CPU_N // 2, "m": m_s // 4, "load_s": 0.0, "compute_s": time.time() - t0,
                            "W": (CPU_N // 2) * (CPU_N // 2 + 1) / 2 * (m_s // 4)})
        kind = "port"
    else:
        kind = "reference"
        with tempfile.TemporaryDirectory(prefix="dgrm_ref_") as wd:
            make_cpu_sample(wd, m_s)
            for _ in range(min(args.warmup, 1)):
                reference_step(wd, threads)
            res = [reference_step(wd, threads) for _ in range(args.steps)]
    t = sum(r["load_s"] + r["compute_s"] for r in res)
    W = sum(r["W"] for r in res)
    val = W / t
    base.update({
        "value": val, "ms_per_step": 1e3 * t / len(res),
        "config": {"workload": f"make-grm {res[-1]['n']}x{m_s if kind == 'reference' else res[-1]['m']} bounded CPU sample "
                               "of configs[1] (20000x100000, 1% missing, MAF filter)",
                   "parallelism": f"1 MPI rank, {threads} OpenMP/BLAS threads"},
        "cpu_baseline": {"value": val, "unit": "pair*SNP/s", "cores": threads, "kind": kind,
                         "sample": sample_desc(res[-1], threads) if kind == "reference" else
                         f"plain-C restatement oracle/grm_oracle.c on n={res[-1]['n']} x m={res[-1]['m']}"},
        "e2e": {"value": val, "unit": "pair*SNP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(base)


def run_gpu(args):
    # torchrun exports OMP_NUM_THREADS=1 to every rank; libdgrm's host threads (they widen the streamed pair counts
    # into the caller's fp64 matrix on the e2e path) would then be one per rank.  Give each rank its share of the
    # box's cores instead; libdgrm's OpenMP runtime reads the variable when the library is loaded, below.
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // int(os.environ["WORLD_SIZE"])))
    import torch
    import torch.distributed as dist
    from cuda.bindings import runtime as cudart
    from dissect_b200 import GrmEngine
    from dissect_b200.engine import PinnedArray
    from dissect_b200.synth import synth_packed_cuda

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the GRM engine has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    G = world
    n = args.n if args.n else int(round(BASE_N * math.sqrt(G)))
    m_file = args.m if args.m else BASE_M

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- synthetic cohort (same seed on every rank -> identical data; rank r keeps its row shard) ----------
    pk_all, maf = synth_packed_cuda(n, m_file, SEED, maf_lo=MAF_LO, maf_hi=MAF_HI, missing=MISSING)
    keep = (maf >= MAF_MIN).cpu().numpy().astype(np.uint8)          # "MAF filter" = --extract list
    m_kept = int(keep.sum())
    W = n * (n + 1) / 2 * m_kept
    bpr = pk_all.shape[1]
    chunk = args.chunk
    # SNP rows are dealt to ranks in contiguous shards of every chunk; the packed chunk is all-gathered
    # over NVLink (the path's one exchange step), then every rank expands it locally.
    # Exchange granularity: `xg` engine chunks travel in one all-gather.  The all-gather kernel cannot run underneath
    # the persistent contraction, so every exchange is a bubble in the MMA stream; fewer, larger
    # exchanges mean fewer bubbles, and all but the first chunk of a group are decoded underneath the contraction.
    xchunk = chunk * max(1, args.xg)
    shard = (xchunk + G - 1) // G
    if G > 1:
        mine = []
        for c0 in range(0, m_file, xchunk):
            r0 = min(c0 + rank * shard, min(c0 + xchunk, m_file))
            r1 = min(r0 + shard, min(c0 + xchunk, m_file))
            mine.append((c0, r0, r1))
        # rows are exchanged with a 16-byte aligned pitch so that the engine's row gather can use 128-bit copies
        bpr_pad = (bpr + 15) // 16 * 16
        send_bufs = [torch.zeros((shard, bpr_pad), dtype=torch.uint8, device="cuda") for _ in range(2)]
        # which of the G*shard gathered rows of a chunk are real and kept (shards are padded to `shard` rows)
        rks = []
        for (c0, r0, r1) in mine:
            c1 = min(c0 + xchunk, m_file)
            rk = np.zeros(G * shard, np.uint8)
            for g in range(G):
                a = c0 + g * shard
                b = min(a + shard, c1)
                if b > a:
                    rk[g * shard: g * shard + (b - a)] = keep[a:b]
            rks.append(rk)
    host_pk = None
    if not args.no_e2e:   # the e2e leg starts from pinned host rows
        host_pk = PinnedArray((m_file, bpr), np.uint8)
        host_pk.array[:] = pk_all.cpu().numpy()
    xch = None
    if G > 1 and args.exchange == "p2p":
        # copy-engine exchange (dissect_b200/exchange.py): shards go straight into every peer's gather buffer
        from dissect_b200.exchange import NBUF, PeerExchange
        xch = PeerExchange(shard, bpr_pad, rank, G, dist.new_group(backend="gloo"))
        if not xch.ok:
            if rank == 0:
                print(f"bench.py: peer exchange unavailable ({xch.error}); using NCCL all-gather", file=sys.stderr)
            xch = None
    if xch is not None:
        pk_src = torch.zeros((m_file, bpr_pad), dtype=torch.uint8, device="cuda")   # contiguous shard sources
        pk_src[:, :bpr] = pk_all
        del pk_all
        torch.cuda.empty_cache()
    elif G > 1:
        gather_bufs = [torch.zeros((G * shard, bpr_pad), dtype=torch.uint8, device="cuda") for _ in range(2)]

    def exchange_and_push(shard_src):
        """shard_src(ci) -> (device address of this rank's contiguous shard rows of chunk ci, row count), readable in
        xch.copy_stream order.  Shards are posted two chunks ahead; everything is stream-ordered (no host wait)."""
        nch, posted = len(mine), 0
        for ci in range(nch):
            while posted < min(nch, ci + NBUF):
                ptr, cnt = shard_src(posted)
                xch.post(ptr, cnt)
                posted += 1
            xch.consume(eng, rks[ci])

    eng = GrmEngine(device=local, rank=rank, world=world, operand=args.operand, count_method=args.count,
                    chunk_snps=chunk, cta_group=args.cg, profile=True, resident_chunks=args.resident,
                    resident_mb=args.resident_mb)
    eng.set_individuals(n)

    def push_device_resident():
        if G == 1:
            eng.push_bed_device(pk_all.data_ptr(), pk_all.stride(0), m_file, row_keep=keep)
            return
        if xch is not None:
            exchange_and_push(lambda ci: (pk_src.data_ptr() + mine[ci][1] * bpr_pad, mine[ci][2] - mine[ci][1]))
            return
        stream = torch.cuda.current_stream().cuda_stream
        for ci, (c0, r0, r1) in enumerate(mine):
            gather_buf, send_buf = gather_bufs[ci & 1], send_bufs[ci & 1]
            if r1 > r0:
                send_buf[: r1 - r0, :bpr].copy_(pk_all[r0:r1])
            dist.all_gather_into_tensor(gather_buf, send_buf)
            eng.push_bed_device(gather_buf.data_ptr(), gather_buf.stride(0), G * shard, row_keep=rks[ci], stream=stream)
            # two exchange buffers: the all-gather of chunk c+2 into this one is ordered after the engine read it
            eng.release_inputs(stream)

    def step_device():
        eng.begin(m_kept)
        push_device_resident()
        eng.finish()

    # ---- value: device-resident inputs ------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    barrier()
    eng.kernel_times()  # reading resets the per-class counters: the timed region starts from zero
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    barrier()
    # device clock: events on this rank's current stream; every dgrm_finish returns only when the engine's own
    # streams are idle, so the closing event cannot be recorded before the last GRM is complete
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step_device()
    ev1.record()
    barrier()
    wall_dt = time.perf_counter() - t0
    dt = ev0.elapsed_time(ev1) * 1e-3
    clocks = sampler.stop()
    kt = eng.kernel_times()
    if world > 1:
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    value = args.steps * W / dt

    # ---- K1 (decode) kernels alone, against the HBM roofline --------------------------------------------
    # In the timed steps above they run underneath the contraction (second stream), so their own event times
    # are inflated by the SMs they do not get; three extra, untimed-for-`value` steps with the overlap switched
    # off give their stand-alone bandwidth (bytes_in + bytes_out over t_K1, peak = measured copy BW).
    k1 = None
    if G == 1:
        os.environ["DGRM_OVERLAP"] = "0"
        eng1 = GrmEngine(device=local, operand=args.operand, count_method=args.count, chunk_snps=chunk,
                         cta_group=args.cg, profile=True, resident_chunks=args.resident, resident_mb=args.resident_mb)
        del os.environ["DGRM_OVERLAP"]
        eng1.set_individuals(n)
        reps = 3
        for it in range(reps + 1):
            eng1.begin(m_kept)
            eng1.push_bed_device(pk_all.data_ptr(), pk_all.stride(0), m_file, row_keep=keep)
            eng1.finish()
            if it == 0:
                eng1.kernel_times()
        kt1 = eng1.kernel_times()
        eng1.close()
        geno = float(n) * m_kept
        n_pad = (n + 255) // 256 * 256
        alg = {"gather": 0.5 * geno,                                   # 1/4 B read + 1/4 B written per genotype
               "count": 0.25 * geno + 72.0 * m_kept,                   # packed rows in, 40 B stats + 32 B table out
               "expand": 0.25 * geno + (2.0 + (0.5 if args.count == "fp4" else 1.0 if args.count == "mma" else 0.125))
               * n_pad * ((m_kept + 255) // 256 * 256)}                # operands are materialised: listed (§8d)
        peaks1 = load_peaks()
        k1 = {"bound": "hbm", "peak": peaks1["hbm_gbs"], "unit": "GB/s", "overlap": "off for this measurement",
              "kernels": {k: {"ms_per_step": kt1[k][0] / reps, "algorithmic_bytes": alg[k],
                              "achieved": alg[k] / (kt1[k][0] / reps * 1e-3) / 1e9,
                              "frac": alg[k] / (kt1[k][0] / reps * 1e-3) / 1e9 / peaks1["hbm_gbs"]}
                          for k in ("gather", "count", "expand")}}

    # ---- e2e: host buffers in, block-cyclic lower triangles of A and N out, every step -----------------
    e2e = None
    if G == 1 and not args.no_e2e:
        A_host = PinnedArray((n, n), np.float64, order="F")
        N_host = PinnedArray((n, n), np.float64, order="F")

        # the caller's Matrix::m buffers are bound once; every dgrm_finish streams the finished tile groups
        # into them while later groups are still on the tensor cores
        eng.bind_output("A", A_host.ptr, n, bs=(256, 256), uplo="L")
        eng.bind_output("N", N_host.ptr, n, bs=(256, 256), uplo="L")

        def step_e2e():
            eng.begin(m_kept)
            eng.push_bed_host(host_pk.ptr, row_keep=keep, pitch=bpr, nrows=m_file)
            eng.finish()

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        ke = max(1, min(args.steps, 3))
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(ke):
            step_e2e()
        ee1.record()
        barrier()
        dte = ee0.elapsed_time(ee1) * 1e-3
        tri = n * (n + 1) // 2
        e2e = {"value": ke * W / dte, "unit": "pair*SNP/s", "ms_per_step": 1e3 * dte / ke,
               "h2d_bytes_per_step": int(m_file * bpr), "d2h_bytes_per_step": int(2 * tri * 8),
               "result": "A and N lower triangles (fp64) into pinned block-cyclic Matrix::m buffers, bs 256, 1x1 grid",
               "check_trace_A": float(np.trace(A_host.array[:2000, :2000]))}
        eng.unbind_outputs()
    elif G > 1 and not args.no_e2e:
        # per-rank H2D of its row shards + all-gather + compute + D2H of the tiles this rank owns
        ntl = eng.num_tiles
        tiles_A = PinnedArray((max(ntl, 1), 256, 128), np.float64)
        tiles_N = PinnedArray((max(ntl, 1), 256, 128), np.float64)

        dbg = os.environ.get("DGRM_BENCH_DEBUG") and rank == 0
        eng.bind_tiles("A", tiles_A.ptr)      # finished tile groups stream out while later groups are contracted
        eng.bind_tiles("N", tiles_N.ptr)

        def h2d_shard(ci, stream):
            # this rank's shard of the chunk: pinned host rows -> device (row pitch padded to 16 bytes)
            c0, r0, r1 = mine[ci]
            send_buf = send_bufs[ci & 1]
            if r1 > r0:
                cudart.cudaMemcpy2DAsync(send_buf.data_ptr(), bpr_pad, host_pk.ptr + r0 * bpr, bpr, bpr, r1 - r0,
                                         cudart.cudaMemcpyKind.cudaMemcpyHostToDevice, stream)
            return send_buf, r1 - r0

        def step_e2e():
            t_a = time.perf_counter()
            eng.begin(m_kept)
            if xch is not None:
                exchange_and_push(lambda ci: (lambda sb_cnt: (sb_cnt[0].data_ptr(), sb_cnt[1]))(h2d_shard(ci, xch.copy_stream.cuda_stream)))
            else:
                stream = torch.cuda.current_stream().cuda_stream
                for ci in range(len(mine)):
                    gather_buf = gather_bufs[ci & 1]
                    send_buf, _ = h2d_shard(ci, stream)
                    dist.all_gather_into_tensor(gather_buf, send_buf)
                    eng.push_bed_device(gather_buf.data_ptr(), gather_buf.stride(0), G * shard, row_keep=rks[ci], stream=stream)
                    eng.release_inputs(stream)
            t_b = time.perf_counter()
            eng.finish()
            t_c = time.perf_counter()
            if dbg:
                print(f"[e2e rank0] push {1e3*(t_b-t_a):.1f} ms, finish (compute + streamed tiles) {1e3*(t_c-t_b):.1f} ms",
                      file=sys.stderr)

        step_e2e()
        barrier()
        ke = max(1, min(args.steps, 2))
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(ke):
            step_e2e()
        ee1.record()
        barrier()
        dte = ee0.elapsed_time(ee1) * 1e-3
        t = torch.tensor([dte], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dte = float(t.item())
        e2e = {"value": ke * W / dte, "unit": "pair*SNP/s", "ms_per_step": 1e3 * dte / ke,
               "h2d_bytes_per_step": int(sum(r1 - r0 for (_, r0, r1) in mine) * bpr),
               "d2h_bytes_per_step": int(2 * ntl * 128 * 256 * 8),
               "result": "per rank: its packed row shards in, its owned 128x256 tiles of A and N (fp64) streamed out "
                         "into bound pinned buffers",
               "check_tile_sum_A": float(tiles_A.array[0, 0, 0])}
        eng.unbind_outputs()

    if xch is not None:
        xch.close()
    if rank != 0:
        eng.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    syrk_ms, syrk_launches = kt["syrk"]
    launches = sum(v[1] for k, v in kt.items() if k != "d2h")
    # algorithmic flops of the contraction: 2 per pair*SNP, lower triangle incl. diagonal;
    # with G ranks each rank's launches cover 1/G of the triangle.
    flops_total = 2.0 * W * args.steps / G
    achieved = flops_total / (syrk_ms * 1e-3) / 1e12 if syrk_ms else None
    peak = peaks["bf16_tflops_sustained"]
    prof_traffic = None
    pj = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(pj):
        prof_traffic = json.load(open(pj)).get("dram_bytes_per_launch")
    out = {
        "metric": "grm_pair_snp_per_s", "value": value, "unit": "pair*SNP/s", "n_gpus": G, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "wall_ms_per_step": 1e3 * wall_dt / args.steps,
        "timing": "CUDA events on the rank's stream around the K steps (each step ends in dgrm_finish, which drains the "
                  "engine's streams), max over ranks", "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": {"fp16": "f16 operands / f32 chunk accum / f64 across chunks; counts exact int32",
                                       "bf16": "bf16 operands / f32 / f64; counts exact int32",
                                       "fp16x2": "f16 hi+lo split (3 MMAs) / f32 / f64"}[args.operand],
        "data": "synthetic",
        "config": {
            "workload": (f"make-grm {n} individuals x {m_file} SNPs ({m_kept} kept by the MAF>=0.01 extract list), "
                         f"1% missing, seed {SEED}" + ("" if G == 1 else f"; weak scaling: n = 20000*sqrt({G})")),
            "pair_snp_per_step": W, "chunk_snps": chunk, "exchange_chunks": max(1, args.xg) if G > 1 else None, "cta_group": args.cg, "operand": args.operand,
            "count_method": args.count,
            "parallelism": "1 GPU" if G == 1 else (
                f"{G} GPUs: lower-triangle tile sets, packed chunks all-gathered "
                + ("by copy-engine peer writes over NVLink (CUDA IPC), gloo host barrier per chunk" if xch is not None
                   else "with NCCL")),
            "l2": "inputs exceed L2 (packed rows + operand chunk + S/N tiles >> 126 MB)",
        },
        "pct_of_bf16_peak": {"burst": 100 * 2 * value / G / (peaks["bf16_tflops"] * 1e12),
                             "sustained": 100 * 2 * value / G / (peaks["bf16_tflops_sustained"] * 1e12),
                             "peaks": peaks["source"]},
        "roofline": {"kernel": "grm_syrk_kernel (tcgen05 kind::f16)", "bound": "tensor", "achieved": achieved,
                     "peak": peak, "peak_kind": f"bf16 sustained, {peaks['source']} (burst {peaks['bf16_tflops']})",
                     "unit": "TFLOP/s", "frac": achieved / peak if achieved else None, "traffic": prof_traffic,
                     "launches": syrk_launches, "avg_launch_ms": syrk_ms / max(syrk_launches, 1)},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
        "kernel_ms_note": "decode kernels (gather/count/expand) run on a second stream underneath the contraction: "
                          "their times here are inflated by the overlap, see roofline_k1 for them alone",
        "roofline_k1": k1,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
    }
    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    if G == 1 and not args.no_cpu:
        try:
            with tempfile.TemporaryDirectory(prefix="dgrm_cpu_") as wd:
                make_cpu_sample(wd)
                threads = cpu_threads()
                r = reference_step(wd, threads)
            v = r["W"] / (r["load_s"] + r["compute_s"])
            out["cpu_baseline"] = {"value": v, "unit": "pair*SNP/s", "cores": threads, "kind": "reference",
                                   "sample": sample_desc(r, threads)}
        except Exception as ex:  # the baseline is reported, never required
            out["cpu_baseline"] = {"value": None, "unit": "pair*SNP/s", "cores": cpu_threads(), "kind": "reference",
                                   "sample": f"unavailable: {ex}"[:300]}
    emit(out)
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line, written with a single syscall to the process's original stdout.  Everything else that
    writes to fd 1 (NCCL prints its version there) has been pointed at stderr by main()."""
    line = (json.dumps(obj) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    off = 0
    while off < len(line):
        off += os.write(fd, line[off:])


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)          # C-level chatter (NCCL, drivers) goes to stderr; stdout carries only the JSON line
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="dgrm", choices=["dgrm", "reference"])
    ap.add_argument("--n", "--individuals", dest="n", type=int, default=0, help="override individuals (default 20000*sqrt(gpus))")
    ap.add_argument("--m", "--snps", dest="m", type=int, default=0, help="override SNPs in the file (default 100000)")
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--xg", type=int, default=1, help="engine chunks per all-gather (multi-GPU exchange granularity)")
    ap.add_argument("--exchange", default="nccl", choices=["p2p", "nccl"],
                    help="multi-GPU chunk exchange: NCCL all-gather (default) or the experimental copy-engine peer "
                         "writes of dissect_b200/exchange.py (correct, but measured slower so far: DESIGN.md §6)")
    ap.add_argument("--cg", type=int, default=2)
    ap.add_argument("--operand", default="fp16", choices=["fp16", "bf16", "fp16x2"])
    ap.add_argument("--count", default="fp4", choices=["mma", "popc", "fp4"])
    ap.add_argument("--resident", type=int, default=0, help="resident chunk slots (0 = auto, 1 = streaming)")
    ap.add_argument("--resident-mb", type=int, default=16384, help="memory budget for resident operands (MiB)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
