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
# bounded CPU sample of the same workload for the reference arm: the SAME 20,000 individuals (the reference's cost
# split between its O(n m) loader and its O(n^2 m) dsyrk depends on n), the SNP count cut to ~10 s of host work
# (measured: the reference spends ~16 ms per SNP at n = 20000 on 8 cores, linear in the SNP count).  The reference
# ARM repeats its sample steps + 1 times, so it uses fewer SNPs than the single cpu_baseline run of the GPU arm.
CPU_N, CPU_M, CPU_M_ARM = 20000, 3000, 1500


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
            f"on a 1x1 BLACS shim, OpenBLAS dsyrk, {threads} threads) on n={r['n']} individuals x m={r['m']} kept SNPs: "
            f"configs[1]'s 20000 individuals with the SNP count cut from 100000 (same "
            f"distributions, 1% missing, MAF>=0.01 extract list): load {r['load_s']:.2f}s + compute "
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
    # every step is the same bounded sample (n = 20000, CPU_M SNPs in the file), whatever K is: ~8 s per step
    runs = args.steps + min(args.warmup, 1)
    m_s = CPU_M_ARM
    if not os.path.exists(exe):
        # the reference-built oracle travels prebuilt; without it fall back to the plain-C port
        from oracle import c_oracle
        from dissect_b200 import bedio
        if not c_oracle.available():
            base.update({"unavailable": "neither oracle/_ref/grmtime nor oracle/_ref/libgrm_oracle.so is built"})
            emit(base)
            return
        pn, pm = 2000, 1000      # the scalar plain-C port is O(n^2 m) on one core: a smaller sample of the family
        pk = bedio.synth_packed(pn, pm, SEED, maf_lo=0.05, maf_hi=MAF_HI, missing=MISSING)
        res = []
        for i in range(runs):
            t0 = time.time()
            c_oracle.grm(pk, pn)
            if i >= runs - args.steps:
                res.append({"n": pn, "m": pm, "load_s": 0.0, "compute_s": time.time() - t0, "W": pn * (pn + 1) / 2 * pm})
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
        "config": {"workload": f"make-grm {res[-1]['n']} individuals x {m_s if kind == 'reference' else res[-1]['m']} SNPs "
                               f"({res[-1]['m']} kept): bounded CPU sample of configs[1] (20000 individuals x 100000 SNPs, 1% "
                               "missing, MAF filter) with the SNP count cut, the same sample every step",
                   "parallelism": f"1 MPI rank, {threads} OpenMP/BLAS threads"},
        "load_s_per_step": sum(r["load_s"] for r in res) / len(res),
        "compute_s_per_step": sum(r["compute_s"] for r in res) / len(res),
        "cpu_baseline": {"value": val, "unit": "pair*SNP/s", "cores": threads, "kind": kind,
                         "load_s": sum(r["load_s"] for r in res) / len(res),
                         "compute_s": sum(r["compute_s"] for r in res) / len(res),
                         "sample": sample_desc(res[-1], threads) if kind == "reference" else
                         f"plain-C restatement oracle/grm_oracle.c on n={res[-1]['n']} x m={res[-1]['m']}"},
        "e2e": {"value": val, "unit": "pair*SNP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    })
    emit(base)


def parity_block(rank, world, local, args, dist):
    """Outside the timed region: the n = 1500 x m = 3000 oracle case through the SAME driver the timed steps use
    (dissect_b200.multi.MultiGpuGrm, same exchange, same operand / count method), tiles gathered on rank 0 and
    compared with the numpy oracle."""
    import torch
    from dissect_b200 import bedio
    from dissect_b200.engine import assemble_from_tiles
    from dissect_b200.multi import DeviceRows, MultiGpuGrm
    n, m, chunk = 1500, 3000, 1024
    pk = bedio.synth_packed(n, m, 17, missing=0.02)         # same seed everywhere; a rank only reads its shards
    mg = MultiGpuGrm(n, rank=rank, world=world, device=local, chunk_snps=chunk, xg=1, exchange=args.exchange,
                     operand=args.operand, count_method=args.count, cta_group=args.cg)
    dev_pk = torch.zeros((m, mg.pitch), dtype=torch.uint8, device="cuda")
    dev_pk[:, : pk.shape[1]] = torch.from_numpy(pk).cuda()
    mg.build(DeviceRows(dev_pk), m)
    st = mg.eng.snp_stats()
    sets = {k: mg.gather_tiles(k) for k in ("A", "N")}
    out = None
    if rank == 0:
        from oracle import grm_oracle as orc
        ref = orc.grm(orc.decode_rows(pk, n))
        low = np.tril(np.ones((n, n), bool))
        N = assemble_from_tiles(n, sets["N"])
        A = assemble_from_tiles(n, sets["A"])
        stats_ok = bool(np.array_equal(st["freq"], ref["stats"]["freq"]) and all(
            np.array_equal(st[k].view(np.uint64), ref["stats"][k].view(np.uint64)) for k in ("p1", "p2", "sd")))
        covered = not (np.isnan(A[low]).any() or np.isnan(N[low]).any())
        eN = float(np.nanmax(np.abs(N[low] - ref["N"][low])))
        eA = float(np.nanmax(np.abs(A[low] - ref["A"][low])))
        tolA = {"fp16": 1e-4, "bf16": 1e-3, "fp16x2": 5e-6}[args.operand]
        out = {"case": f"n={n} m={m} chunk={chunk}, 2% missing, seed 17, {world} rank(s), exchange {mg.exchange}",
               "oracle": "oracle/grm_oracle.py (numpy restatement pinned to the reference binary)",
               "N_exact": covered and eN == 0.0, "A_max_err": eA, "A_tol": tolA, "stats_bitexact": stats_ok,
               "triangle_covered_once": covered, "ok": bool(covered and eN == 0.0 and eA <= tolA and stats_ok)}
    mg.close()
    del dev_pk
    torch.cuda.empty_cache()
    return out


CONFIGS = {
    # name: individuals, SNPs in the file, seed, SNP groups (BASELINE.json configs[1..4])
    "c2": dict(n=20000, m=100000, seed=2, groups=0),
    "c3": dict(n=100000, m=500000, seed=3, groups=0),
    "c4": dict(n=50000, m=300000, seed=4, groups=22),
    "c5": dict(n=50000, m=200000, seed=5, groups=0),
}


def run_gpu(args):
    # torchrun exports OMP_NUM_THREADS=1 to every rank; libdgrm's host threads (they widen the streamed pair counts
    # into the caller's fp64 matrix on the e2e path) would then be one per rank.  Give each rank its share of the
    # box's cores instead; libdgrm's OpenMP runtime reads the variable when the library is loaded, below.
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        os.environ["OMP_NUM_THREADS"] = str(max(1, cpu_threads() // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))
    import torch
    import torch.distributed as dist
    from dissect_b200 import GrmEngine
    from dissect_b200.engine import PinnedArray
    from dissect_b200.multi import DeviceRows, HostRows, MultiGpuGrm
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
    cfgd = CONFIGS[args.config]
    weak = args.config == "c2" and not args.n
    n = args.n if args.n else (int(round(BASE_N * math.sqrt(G))) if weak else cfgd["n"])
    m_file = args.m if args.m else cfgd["m"]
    seed = cfgd["seed"]
    if cfgd["groups"] and G > 1:
        raise SystemExit("bench.py: --config c4 (SNP-group GRMs from one resident cohort) runs on one GPU")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    chunk = args.chunk
    # Transport of the streamed outputs (e2e leg).  Narrow transports (A as fp32, N as 16-bit deficiencies) cut the PCIe
    # bytes but are widened to the caller's doubles by host threads; with only a few host cores per rank that widening is
    # the bottleneck (measured r2, 2 ranks on 24 vCPUs), so then the copy engine delivers doubles directly.
    threads_per_rank = max(1, cpu_threads() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", world))))
    a_tr = args.a_transport if args.a_transport != "auto" else ("fp32" if threads_per_rank >= 16 else "fp64")
    n_tr = args.n_transport if args.n_transport != "auto" else ("narrow" if threads_per_rank >= 2 else "fp64")
    mg = MultiGpuGrm(n, rank=rank, world=world, device=local, chunk_snps=chunk, xg=args.xg, exchange=args.exchange,
                     operand=args.operand, count_method=args.count, cta_group=args.cg, profile=True,
                     resident_chunks=args.resident, resident_mb=args.resident_mb, a_transport=a_tr, n_transport=n_tr)
    eng = mg.eng
    # ---- synthetic cohort (same seed on every rank -> identical data; a rank only READS its own row shards) ----
    pk_raw, maf = synth_packed_cuda(n, m_file, seed, maf_lo=MAF_LO, maf_hi=MAF_HI, missing=MISSING)
    keep = (maf >= MAF_MIN).cpu().numpy().astype(np.uint8)          # "MAF filter" = --extract list
    m_kept = int(keep.sum())
    bpr = pk_raw.shape[1]
    if G > 1 and mg.pitch != bpr:     # rows at the exchange pitch: shards go out without staging
        pk_all = torch.zeros((m_file, mg.pitch), dtype=torch.uint8, device="cuda")
        pk_all[:, :bpr] = pk_raw
        del pk_raw
        torch.cuda.empty_cache()
    else:
        pk_all = pk_raw
    W = n * (n + 1) / 2 * m_kept
    plan = mg.prepare(m_file, keep)
    host_pk = None
    if not args.no_e2e:   # the e2e leg starts from pinned host rows
        host_pk = PinnedArray((m_file, bpr), np.uint8)
        host_pk.array[:] = pk_all[:, :bpr].cpu().numpy()
    dev_src = DeviceRows(pk_all)

    groups = None
    if cfgd["groups"]:
        # C4: 22 SNP groups with sizes proportional to the human chromosome lengths
        chr_len = np.array([248, 242, 198, 190, 181, 171, 159, 145, 138, 133, 135, 133, 114, 107, 102, 90, 83, 80, 58, 64,
                            46, 50], dtype=np.float64)[: cfgd["groups"]]
        edges = np.concatenate([[0], np.round(np.cumsum(chr_len) / chr_len.sum() * m_file).astype(np.int64)])
        groups = [(int(edges[g]), int(edges[g + 1])) for g in range(len(chr_len))]
        eng.store_bed_device(pk_all.data_ptr(), pk_all.stride(0), m_file)
        W = W * 2      # the global GRM plus every group GRM again for its leave-one-chromosome-out view

    def step_device():
        if groups is None:
            mg.build(dev_src, m_file, keep)
            return
        # global GRM from the resident cohort, kept as (S, N); then per chromosome: group GRM and its LOCO view
        eng.begin(m_kept)
        eng.push_stored(0, m_file, row_keep=keep)
        eng.finish()
        eng.save("copy")
        for (g0, g1) in groups:
            eng.begin(max(1, int(keep[g0:g1].sum())))
            eng.push_stored(g0, g1 - g0, row_keep=keep[g0:g1])
            eng.finish()
            eng.set_view("saved-current")        # global - chromosome (singlereml.cpp:298)
            eng.low_overlap_pairs(10.0, cap=16)  # a reader of the LOCO view: sanitizeKernel's scan over (S, N)
            eng.set_view("current")

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
    per_rank = None
    if world > 1:
        # every rank's own device time and kernel-class times (imbalance is visible in the line), then the max
        mine_t = torch.tensor([dt * 1e3 / args.steps, kt["syrk"][0] / args.steps, kt["paircnt"][0] / args.steps,
                               kt["expand"][0] / args.steps, kt["gather"][0] / args.steps, float(len(eng.tile_coords()))],
                              dtype=torch.float64, device="cuda")
        allt = [torch.zeros_like(mine_t) for _ in range(world)]
        dist.all_gather(allt, mine_t)
        per_rank = {"ms_per_step": [round(float(x[0]), 3) for x in allt], "syrk_ms": [round(float(x[1]), 3) for x in allt],
                    "paircnt_ms": [round(float(x[2]), 3) for x in allt], "expand_ms": [round(float(x[3]), 3) for x in allt],
                    "decode_ms": [round(float(x[4]), 3) for x in allt], "tiles": [int(x[5]) for x in allt]}
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    value = args.steps * W / dt

    # ---- K1 (decode) kernels alone, against the HBM roofline --------------------------------------------
    # In the timed steps above they run underneath the contraction (second stream), so their own event times
    # are inflated by the SMs they do not get; three extra, untimed-for-`value` steps with the overlap switched
    # off give their stand-alone bandwidth (bytes_in + bytes_out over t_K1, peak = measured copy BW).
    k1 = None
    if G == 1 and groups is None and not args.no_k1:
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
        qb = 0.5 if args.count == "fp4" else 1.0 if args.count == "mma" else 0.125
        alg = {"decode": 0.25 * geno + 0.25 * geno + 72.0 * m_kept,     # packed rows in, aligned chunk + 40 B stats + 32 B table out
               "expand": 0.25 * geno + (2.0 * (2 if args.operand == "fp16x2" else 1) + qb)
               * n_pad * ((m_kept + 255) // 256 * 256)}                # operands are materialised: listed (§8d)
        names = {"decode": ("gather", "count"), "expand": ("expand",)}
        peaks1 = load_peaks()
        # write-only stream in the same run (cudaMemset of 4 GiB): bed_expand writes 10x what it reads, and a pure write
        # stream does not reach the read+write figure that MEASURED_PEAKS.json's copy test gives
        wbuf = torch.empty(4 << 30, dtype=torch.uint8, device="cuda")
        wbuf.zero_()
        torch.cuda.synchronize()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(3):
            wbuf.zero_()
        w1.record()
        torch.cuda.synchronize()
        write_only_gbs = 3 * wbuf.numel() / (w0.elapsed_time(w1) * 1e-3) / 1e9
        del wbuf
        kern = {}
        for k, parts in names.items():
            ms = sum(kt1[p][0] for p in parts) / reps
            if ms > 0:
                kern[k] = {"ms_per_step": ms, "algorithmic_bytes": alg[k], "achieved": alg[k] / (ms * 1e-3) / 1e9,
                           "frac": alg[k] / (ms * 1e-3) / 1e9 / peaks1["hbm_gbs"],
                           "launches_per_step": sum(kt1[p][1] for p in parts) / reps}
        if "expand" in kern:
            kern["expand"]["frac_of_write_only_stream"] = kern["expand"]["achieved"] / write_only_gbs
        k1 = {"bound": "hbm", "peak": peaks1["hbm_gbs"], "unit": "GB/s", "overlap": "off for this measurement",
              "write_only_stream_gbs": write_only_gbs,
              "write_only_note": "cudaMemset of 4 GiB timed in this run; `peak` is the copy (read + write) figure of MEASURED_PEAKS.json",
              "kernels": kern}

    # ---- e2e: host buffers in, A and N (lower triangle / this rank's tiles) out, every step ------------------
    e2e = None
    tiles_A = tiles_N = None
    if groups is not None:
        pass
    elif G == 1 and not args.no_e2e:
        A_host = PinnedArray((n, n), np.float64, order="F")
        N_host = PinnedArray((n, n), np.float64, order="F")
        # the caller's Matrix::m buffers are bound once; every dgrm_finish streams the finished tile groups
        # into them while later groups are still on the tensor cores
        eng.bind_output("A", A_host.ptr, n, bs=(256, 256), uplo="L")
        eng.bind_output("N", N_host.ptr, n, bs=(256, 256), uplo="L")
        host_src = HostRows(host_pk.ptr, bpr)

        def step_e2e():
            mg.build(host_src, m_file, keep)

        for _ in range(max(1, min(args.warmup, 2))):
            step_e2e()
        barrier()
        ke = max(1, args.steps)
        ee0, ee1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ee0.record()
        for _ in range(ke):
            step_e2e()
        ee1.record()
        barrier()
        dte = ee0.elapsed_time(ee1) * 1e-3
        e2e = {"value": ke * W / dte, "unit": "pair*SNP/s", "ms_per_step": 1e3 * dte / ke, "steps": ke,
               "h2d_bytes_per_step": int(m_file * bpr), "d2h_bytes_per_step": int(eng.last_d2h_bytes),
               "result": "A and N lower triangles as fp64 in pinned block-cyclic Matrix::m buffers (bs 256, 1x1 grid); "
                         "d2h_bytes_per_step = bytes the copy engine really moved (A as " + a_tr +
                         (", N as 16-bit deficiencies m - N or int32, widened to fp64 by host threads)" if n_tr == "narrow" else ", N as fp64)"),
               "host_threads_per_rank": threads_per_rank,
               "check_trace_A": float(np.trace(A_host.array[:2000, :2000]))}
        eng.unbind_outputs()
        del A_host, N_host
    elif G > 1 and not args.no_e2e:
        # per-rank H2D of its row shards + all-gather + compute + D2H of the tiles this rank owns
        ntl = eng.num_tiles
        tiles_A = PinnedArray((max(ntl, 1), 256, 128), np.float64)
        tiles_N = PinnedArray((max(ntl, 1), 256, 128), np.float64)
        eng.bind_tiles("A", tiles_A.ptr)      # finished tile groups stream out while later groups are contracted
        eng.bind_tiles("N", tiles_N.ptr)
        host_src = HostRows(host_pk.ptr, bpr)

        def step_e2e():
            mg.build(host_src, m_file, keep)

        step_e2e()
        barrier()
        ke = max(1, args.steps)
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
        e2e = {"value": ke * W / dte, "unit": "pair*SNP/s", "ms_per_step": 1e3 * dte / ke, "steps": ke,
               "h2d_bytes_per_step": int(plan.rows_of(rank) * bpr), "d2h_bytes_per_step": int(eng.last_d2h_bytes),
               "bytes_are": "this rank's; every rank moves about the same",
               "result": "per rank: its packed row shards in, its owned 128x256 tiles of A and N as fp64 in bound pinned "
                         "buffers; d2h_bytes_per_step = bytes the copy engine really moved (A as " + a_tr +
                         (", N as 16-bit deficiencies or int32, widened by host threads)" if n_tr == "narrow" else ", N as fp64)"),
               "host_threads_per_rank": threads_per_rank}
        eng.unbind_outputs()

    # ---- e2e against the host link: bytes that really crossed / the measured concurrent-copy ceiling --------------
    if e2e is not None:
        pj = os.path.join(ROOT, "profiles", "r2_d2h_concurrency_n8.json")
        if os.path.exists(pj):
            rows = [r for r in json.load(open(pj))["rows"] if r["direction"] == "d2h"]
            near = min(rows, key=lambda r: abs(r["ranks_active"] - G))
            tot = torch.tensor([float(e2e["d2h_bytes_per_step"])], dtype=torch.float64, device="cuda")
            if world > 1:
                dist.all_reduce(tot, op=dist.ReduceOp.SUM)
            floor_ms = float(tot.item()) / (near["aggregate_GBps"] * 1e9) * 1e3
            e2e["roofline_e2e"] = {
                "bound": "host link, device->host", "d2h_bytes_all_ranks": float(tot.item()),
                "ceiling_gbs": near["aggregate_GBps"], "ceiling_ranks": near["ranks_active"],
                "ceiling_source": "profiles/r2_d2h_concurrency_n8.json: plain pinned cudaMemcpy D2H issued concurrently on k ranks of "
                                  "an 8-GPU box of this pool (56 / 69 / 72 / 93 GB/s aggregate at 1 / 2 / 4 / 8 ranks: the GPUs share the link)",
                "floor_ms": floor_ms, "frac": floor_ms / e2e["ms_per_step"],
                "note": "floor = D2H bytes of all ranks / ceiling; the step also has to compute (ms_per_step of `value`) and upload"}

    # ---- correctness of what was just timed (outside the timed regions) --------------------------------------
    checks = None
    if not args.no_check and groups is None:
        from tools import fullsize_checks
        if tiles_A is None:       # no e2e leg (or one GPU): take the tiles of the last timed GRM, slab by slab
            tA = lambda t0, t1: eng.fetch_tiles_range("A", t0, t1 - t0)      # noqa: E731
            tN = lambda t0, t1: eng.fetch_tiles_range("N", t0, t1 - t0)      # noqa: E731
        else:
            tA, tN = tiles_A.array, tiles_N.array
        checks = fullsize_checks.run(pk_all, keep, n, m_kept, eng.tile_coords(), tA, tN, eng.snp_stats(), rank, world,
                                     torch.device("cuda", local), operand=args.operand, dist=dist)
        del tA, tN
    tiles_A = tiles_N = None
    del pk_all
    dev_src = None
    torch.cuda.empty_cache()
    parity = None if args.no_check else parity_block(rank, world, local, args, dist)

    if rank != 0:
        mg.close()
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    syrk_ms, syrk_launches = kt["syrk"]
    pc_ms, pc_launches = kt["paircnt"]
    launches = sum(v[1] for k, v in kt.items() if k != "d2h")
    # algorithmic flops of the contraction: 2 per pair*SNP, lower triangle incl. diagonal;
    # with G ranks each rank's launches cover 1/G of the triangle.
    flops_total = 2.0 * W * args.steps / G
    achieved = flops_total / (syrk_ms * 1e-3) / 1e12 if syrk_ms else None
    peak = peaks["bf16_tflops_sustained"]
    prof_traffic = traffic_src = None
    pj = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(pj) and G == 1 and args.config == "c2" and not args.n and not args.m:
        tj = json.load(open(pj))
        if tj.get("chunk_snps", 16384) == chunk and tj.get("operand", "fp16") == args.operand:
            prof_traffic = tj.get("dram_bytes_per_launch")
            traffic_src = tj.get("source", "profiles/syrk_traffic.json") + " (one ncu --set full capture of this command's kernel, not measured in this run)"
    out = {
        "metric": "grm_pair_snp_per_s", "value": value, "unit": "pair*SNP/s", "n_gpus": G, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "wall_ms_per_step": 1e3 * wall_dt / args.steps,
        "timing": "CUDA events on the rank's stream around the K steps (each step ends in dgrm_finish, which drains the "
                  "engine's streams), max over ranks", "higher_is_better": True, "scaling": "weak" if weak or G == 1 else "strong",
        "vs_baseline": None, "dtype": {"fp16": "f16 operands / f32 chunk accum / f64 across chunks; counts exact int32",
                                       "bf16": "bf16 operands / f32 / f64; counts exact int32",
                                       "fp16x2": "f16 hi+lo split (3 MMAs) / f32 / f64"}[args.operand],
        "data": "synthetic",
        "config": {
            "workload": (f"make-grm {n} individuals x {m_file} SNPs ({m_kept} kept by the MAF>=0.01 extract list), "
                         f"1% missing, seed {seed} (BASELINE configs[{int(args.config[1]) - 1}])"
                         + (f"; weak scaling: n = 20000*sqrt({G})" if weak and G > 1 else "")
                         + (f"; global GRM + {len(groups)} chromosome GRMs and their LOCO views from one resident cohort" if groups else "")),
            "pair_snp_per_step": W, "chunk_snps": chunk, "exchange_chunks": mg.xg if G > 1 else None, "cta_group": args.cg,
            "operand": args.operand, "count_method": args.count,
            "parallelism": "1 GPU" if G == 1 else (
                f"{G} GPUs: lower-triangle tile sets, packed chunks all-gathered "
                + ("by copy-engine peer writes over NVLink (torch symmetric memory), stream-ordered, no SM used" if mg.exchange == "p2p" else "with NCCL all_gather_into_tensor")
                + " (dissect_b200/multi.py)"),
            "l2": "inputs exceed L2 (packed rows + operand chunk + S/N tiles >> 126 MB)",
        },
        "pct_of_bf16_peak": {"burst": 100 * 2 * value / G / (peaks["bf16_tflops"] * 1e12),
                             "sustained": 100 * 2 * value / G / (peaks["bf16_tflops_sustained"] * 1e12),
                             "peaks": peaks["source"]},
        "roofline": {"kernel": "grm_syrk_kernel (tcgen05 kind::f16)", "bound": "tensor", "achieved": achieved,
                     "peak": peak, "peak_kind": f"bf16 sustained, {peaks['source']} (burst {peaks['bf16_tflops']})",
                     "unit": "TFLOP/s", "frac": achieved / peak if achieved else None, "traffic": prof_traffic,
                     "traffic_source": traffic_src,
                     "launches": syrk_launches, "avg_launch_ms": syrk_ms / max(syrk_launches, 1)},
        "roofline_k4": None if not pc_ms else {
            "kernel": {"fp4": "grm_syrk_kernel<.,2,int> (tcgen05 kind::mxf4, e2m1 0/1 operands)",
                       "mma": "grm_syrk_kernel<.,1,int> (tcgen05 kind::f8f6f4)", "popc": "pair_count_popc_kernel"}[args.count],
            "bound": "tensor", "unit": "pair*SNP/s", "achieved": W * args.steps / G / (pc_ms * 1e-3),
            "peak": 4 * peaks["bf16_tflops_sustained"] * 1e12 / 2 if args.count == "fp4" else
                    2 * peaks["bf16_tflops_sustained"] * 1e12 / 2 if args.count == "mma" else None,
            "peak_kind": "dense fp4 (fp8) MMA rate taken as 4x (2x) the measured sustained bf16 rate at 2 flop per "
                         "pair*SNP: nominal ratio, no fp4 GEMM measurement exists in MEASURED_PEAKS.json",
            "launches": pc_launches, "avg_launch_ms": pc_ms / max(pc_launches, 1)},
        "kernel_ms_per_step": {k: v[0] / args.steps for k, v in kt.items()},
        "per_rank": per_rank,
        "kernel_ms_note": "decode kernels (gather/count/expand) run on a second stream underneath the contraction: "
                          "their times here are inflated by the overlap, see roofline_k1 for them alone",
        "roofline_k1": k1,
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": e2e,
        "parity": parity,
        "fullsize_checks": checks,
    }
    if out["roofline_k4"] and out["roofline_k4"]["peak"]:
        out["roofline_k4"]["frac"] = out["roofline_k4"]["achieved"] / out["roofline_k4"]["peak"]
    # ---- CPU baseline beside it (rank 0, N = 1 only) ---------------------------------------------------
    if G == 1 and not args.no_cpu:
        try:
            with tempfile.TemporaryDirectory(prefix="dgrm_cpu_") as wd:
                make_cpu_sample(wd)
                threads = cpu_threads()
                r = reference_step(wd, threads)
            v = r["W"] / (r["load_s"] + r["compute_s"])
            out["cpu_baseline"] = {"value": v, "unit": "pair*SNP/s", "cores": threads, "kind": "reference",
                                   "load_s": r["load_s"], "compute_s": r["compute_s"],
                                   "value_compute_only": r["W"] / r["compute_s"] if r["compute_s"] else None,
                                   "sample": sample_desc(r, threads)}
        except Exception as ex:  # the baseline is reported, never required
            out["cpu_baseline"] = {"value": None, "unit": "pair*SNP/s", "cores": cpu_threads(), "kind": "reference",
                                   "sample": f"unavailable: {ex}"[:300]}
    emit(out)
    mg.close()
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
    ap.add_argument("--config", default="c2", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: c2 = configs[1] (the default, the one the metric is quoted on; individuals "
                         "scale with sqrt(gpus)), c3 / c5 = configs[2] / configs[4] at their full sizes, c4 = configs[3] "
                         "(22 chromosome GRMs + LOCO views from one resident cohort, one GPU)")
    ap.add_argument("--chunk", type=int, default=16384)
    ap.add_argument("--xg", type=int, default=0, help="engine chunks per all-gather (multi-GPU exchange granularity); "
                                                      "0 = as many as a 2 GB exchange-buffer budget allows")
    ap.add_argument("--a-transport", default="auto", choices=["auto", "fp64", "fp32"],
                    help="how streamed A values cross PCIe on the e2e path: fp32 (+ the exact fp64 diagonal; adds <= 6e-8 "
                         "relative, the caller still receives fp64), fp64, or auto = fp32 when a rank has >= 16 host threads "
                         "for the widening, else fp64 (r2: 1 rank x 16 threads 56 vs 69 ms; 2 ranks x 12 threads 86 vs 73 ms)")
    ap.add_argument("--n-transport", default="auto", choices=["auto", "narrow", "fp64"],
                    help="how streamed pair counts cross PCIe: narrow (16-bit deficiencies / int32, widened by host threads), "
                         "fp64 (doubles by the copy engine, no host work), auto = narrow unless a rank has a single host thread (r2, 8 ranks x 4 "
                         "threads: the shared host link, 93 GB/s for all 8 GPUs, is the bound, so fewer bytes win)")
    ap.add_argument("--no-k1", action="store_true", help="skip the stand-alone decode-kernel measurement")
    ap.add_argument("--no-check", action="store_true", help="skip the parity block and the full-size checks")
    ap.add_argument("--exchange", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU chunk exchange: copy-engine peer writes over NVLink through torch symmetric memory "
                         "(p2p: SM-free, runs underneath the contraction; r2: 38.2 vs 40.1 ms per step at 2 GPUs), NCCL "
                         "all-gather (nccl), or auto = p2p when the peer mapping works on this box, else nccl")
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
