"""Multi-GPU parity check, run under torchrun with one rank per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tools/multi_gpu_check.py [--exchange nccl|p2p] [--individuals 1500 --snps 3000 --chunk 1024] [--out DIR]

Drives dissect_b200.multi.MultiGpuGrm — the same exchange loop bench.py times — and checks, against the numpy
oracle: per-SNP statistics bit-exact, N exact and A within tolerance over the whole lower triangle (tiles gathered
on rank 0), the block-cyclic egress over a 2-D process grid (reduce_block_cyclic) and the parallel .grm.dat writer.
Prints one PASS/FAIL line per check and exits non-zero on any failure."""
import argparse
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import bedio  # noqa: E402
from dissect_b200.engine import assemble_from_tiles, numroc  # noqa: E402
from dissect_b200.multi import DeviceRows, HostRows, MultiGpuGrm  # noqa: E402
from dissect_b200.partition import tiles_of_rank  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exchange", default=os.environ.get("DGRM_EXCHANGE", "nccl"), choices=["nccl", "p2p"])
    ap.add_argument("--individuals", "--n", dest="n", type=int, default=1500)   # (torchrun rejects a bare --n as ambiguous)
    ap.add_argument("--snps", "--m", dest="m", type=int, default=3000)
    ap.add_argument("--chunk", type=int, default=1024)
    ap.add_argument("--xg", type=int, default=1)
    ap.add_argument("--out", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, chunk = args.n, args.m, args.chunk
    pk = bedio.synth_packed(n, m, 17, missing=0.02)     # same seed everywhere; each rank only *reads* its shards
    keep = (np.arange(m) % 7 != 3).astype(np.uint8)     # an --extract list that cuts through the shards
    mg = MultiGpuGrm(n, rank=rank, world=world, device=local, chunk_snps=chunk, xg=args.xg, exchange=args.exchange)
    dev_pk = torch.zeros((m, mg.pitch), dtype=torch.uint8, device="cuda")
    dev_pk[:, : pk.shape[1]] = torch.from_numpy(pk).cuda()
    results = []

    def report(name, ok, detail=""):
        results.append(bool(ok))
        if rank == 0:
            print(f"multi_gpu_check world={world} exchange={args.exchange}: {name}: {detail} -> {'PASS' if ok else 'FAIL'}", flush=True)

    ref = None
    low = np.tril(np.ones((n, n), bool))
    if rank == 0:
        from oracle import grm_oracle as orc
        ref = orc.grm(orc.decode_rows(pk[keep != 0], n))

    for leg, source in (("device rows", DeviceRows(dev_pk)), ("host rows", HostRows(pk.ctypes.data, pk.strides[0]))):
        for rep in range(2):          # twice: buffers and sequence numbers are reused across GRMs
            mg.build(source, m, keep)
        assert sorted(map(tuple, mg.eng.tile_coords().tolist())) == sorted(tiles_of_rank(n, rank, world))   # host mirror == engine
        st = mg.eng.snp_stats()
        sets = {k: mg.gather_tiles(k) for k in ("A", "N")}
        if rank == 0:
            N = assemble_from_tiles(n, sets["N"])
            A = assemble_from_tiles(n, sets["A"])
            covered = not (np.isnan(A[low]).any() or np.isnan(N[low]).any())
            eN = float(np.nanmax(np.abs(N[low] - ref["N"][low])))
            eA = float(np.nanmax(np.abs(A[low] - ref["A"][low])))
            stats_ok = np.array_equal(st["freq"], ref["stats"]["freq"]) and all(
                np.array_equal(st[k].view(np.uint64), ref["stats"][k].view(np.uint64)) for k in ("p1", "p2", "sd"))
            report(f"tiles vs oracle ({leg})", covered and eN == 0.0 and eA < 1e-4 and stats_ok,
                   f"N max err {eN}, A max err {eA:.2e}, stats bit-exact {stats_ok}, triangle covered {covered}")
        else:
            report(f"tiles vs oracle ({leg})", True)

    # ---- block-cyclic egress over an R x C grid of the ranks (matrix.cpp:435-487 layout) ------------------------
    R = 2 if world % 2 == 0 else 1
    C = world // R
    for bs in ((64, 64), (100, 36)):
        pr, pc = rank // C, rank % C
        for which in ("A", "N"):
            loc = mg.reduce_block_cyclic(which, bs, (R, C), uplo="L", panel_cols=97)
            gath = [None] * world if rank == 0 else None
            dist.gather_object(loc, gath, dst=0)
            if rank == 0:
                full = np.full((n, n), np.nan)
                for d in range(world):
                    dr, dc = d // C, d % C
                    nr, nc = numroc(n, bs[0], dr, R), numroc(n, bs[1], dc, C)
                    gr = ((np.arange(nr) // bs[0]) * R + dr) * bs[0] + np.arange(nr) % bs[0]
                    gc = ((np.arange(nc) // bs[1]) * C + dc) * bs[1] + np.arange(nc) % bs[1]
                    full[np.ix_(gr, gc)] = gath[d]
                want = assemble_from_tiles(n, sets[which])
                ok = np.array_equal(full[low], want[low]) and np.isnan(full[~low]).all()
                report(f"block-cyclic egress {which} grid {R}x{C} bs {bs}", ok, "lower triangle equal to the tiles, upper untouched")
            else:
                report("egress", True)

    # ---- parallel .grm.dat writer -------------------------------------------------------------------------------
    path = [None]
    if rank == 0:
        d = args.out or tempfile.mkdtemp(prefix="dgrm_mg_")
        os.makedirs(d, exist_ok=True)
        path[0] = os.path.join(d, f"mg{world}.grm.dat")
    dist.broadcast_object_list(path, src=0)
    mg.write_grm_dat(path[0])
    if rank == 0:
        from oracle import grm_oracle as orc
        blob = open(path[0], "rb").read()
        A2, N2 = orc.unpack_grm_dat(blob, n)
        A = assemble_from_tiles(n, sets["A"])
        ok = len(blob) == 14 + (n + 1) * n * 8 and np.array_equal(N2[low], ref["N"][low]) and np.array_equal(A2[low], A[low])
        report("parallel .grm.dat writer", ok, f"{len(blob)} bytes, N block exact, A block equal to the tiles")
    else:
        report("writer", True)

    mg.close()
    ok = all(results)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"multi_gpu_check world={world} exchange={args.exchange}: {'ALL PASS' if t.item() else 'FAILED'}", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if t.item() else 1)


if __name__ == "__main__":
    main()
