"""Multi-GPU parity check (run under torchrun, one rank per GPU): every rank contracts its tile set from
packed rows all-gathered over NCCL; rank 0 merges the tile sets and compares with the oracle.
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dissect_b200 import GrmEngine, bedio  # noqa: E402
from dissect_b200.engine import assemble_from_tiles  # noqa: E402
from dissect_b200.partition import row_shards, tiles_of_rank  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, m, chunk = 1500, 3000, 1024
    pk = bedio.synth_packed(n, m, 17, missing=0.02)     # same seed everywhere; each rank only *uses* its shard
    bpr = pk.shape[1]
    shard = (chunk + world - 1) // world
    eng = GrmEngine(device=local, rank=rank, world=world, chunk_snps=chunk)
    eng.set_individuals(n)
    eng.begin(m)
    mode = os.environ.get("DGRM_EXCHANGE", "nccl")
    stream = torch.cuda.current_stream().cuda_stream
    shards = row_shards(m, chunk, world)

    def row_keep(ranges):
        rk = np.zeros(world * shard, np.uint8)
        for r, (ra, rb) in enumerate(ranges):
            rk[r * shard: r * shard + (rb - ra)] = 1
        return rk

    if mode == "p2p":
        from dissect_b200.exchange import NBUF, PeerExchange
        pitch = (bpr + 15) // 16 * 16
        xch = PeerExchange(shard, pitch, rank, world, dist.new_group(backend="gloo"))
        src = torch.zeros((m, pitch), dtype=torch.uint8, device="cuda")
        src[:, :bpr] = torch.from_numpy(pk).cuda()

        posted = 0
        for ci, (c0, ranges) in enumerate(shards):
            while posted < min(len(shards), ci + NBUF):
                a, b = shards[posted][1][rank]
                xch.post(src.data_ptr() + a * pitch, b - a)
                posted += 1
            xch.consume(eng, row_keep(ranges))
    else:
        gather_buf = torch.zeros((world * shard, bpr), dtype=torch.uint8, device="cuda")
        send = torch.zeros((shard, bpr), dtype=torch.uint8, device="cuda")
        for c0, ranges in shards:
            a, b = ranges[rank]
            send.zero_()
            if b > a:
                send[: b - a].copy_(torch.from_numpy(pk[a:b]))
            dist.all_gather_into_tensor(gather_buf, send)
            eng.push_bed_device(gather_buf.data_ptr(), gather_buf.stride(0), world * shard, row_keep=row_keep(ranges),
                                stream=stream)
            eng.sync()
    eng.finish()
    rc = eng.tile_coords()
    assert sorted(map(tuple, rc.tolist())) == sorted(tiles_of_rank(n, rank, world))   # host mirror == engine
    st = eng.snp_stats()
    payload = {k: (rc, eng.fetch_tiles(k)) for k in ("A", "N")}
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(payload, gathered, dst=0)
    ok = True
    if rank == 0:
        from oracle import grm_oracle as orc
        ref = orc.grm(orc.decode_rows(pk, n))
        low = np.tril(np.ones((n, n), bool))
        N = assemble_from_tiles(n, [g["N"] for g in gathered])
        A = assemble_from_tiles(n, [g["A"] for g in gathered])
        eN = float(np.max(np.abs(N[low] - ref["N"][low])))
        eA = float(np.max(np.abs(A[low] - ref["A"][low])))
        stats_ok = np.array_equal(st["freq"], ref["stats"]["freq"]) and np.array_equal(
            st["sd"].view(np.uint64), ref["stats"]["sd"].view(np.uint64))
        ok = eN == 0.0 and eA < 1e-4 and stats_ok and not np.isnan(A[low]).any()
        print(f"multi_gpu_check world={world} exchange={mode}: N max err {eN}, A max err {eA:.2e}, stats bit-exact {stats_ok} -> "
              f"{'PASS' if ok else 'FAIL'}")
    if mode == "p2p":
        xch.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
