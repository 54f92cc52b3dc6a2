#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29571 bench.py --gpus 4 --steps 10 --warmup 3 > $O/r2_bench_n4_final.json 2> $O/r2_bench_n4_final.err; echo "bench n4 rc $?"
timeout 900 $TR --master-port 29572 bench.py --gpus 4 --steps 10 --warmup 3 --exchange nccl --no-e2e --no-check > $O/r2_bench_n4_nccl.json 2> $O/r2_bench_n4_nccl.err; echo "bench n4 nccl rc $?"
for f in $O/r2_bench_n4_final.json $O/r2_bench_n4_nccl.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
e=d["e2e"]
print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "e2e", e and round(e["ms_per_step"],1), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"])
print("   per rank", d["per_rank"])
PY
done
