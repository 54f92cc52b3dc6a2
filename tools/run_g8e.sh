#!/bin/bash
# 8-GPU session E: final code — p2p parity at 8 ranks, driver-style weak-scaling line, C3
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29601 tools/multi_gpu_check.py --exchange p2p > $O/r2_g8e_check.log 2>&1; grep multi_gpu_check $O/r2_g8e_check.log | tail -9
timeout 900 $TR --master-port 29602 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2_bench_n8_final2.json 2> $O/r2_bench_n8_final2.err; echo "bench n8 rc $?"
timeout 1200 $TR --master-port 29603 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e > $O/r2_bench_c3_n8_final2.json 2> $O/r2_bench_c3_n8_final2.err; echo "bench c3 rc $?"
for f in $O/r2_bench_n8_final2.json $O/r2_bench_c3_n8_final2.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d["e2e"]
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "e2e", e and round(e["ms_per_step"],1), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"], "xg", d["config"]["exchange_chunks"], d["per_rank"]["ms_per_step"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
