#!/bin/bash
# 4-GPU session: the compact (macro-block) partition at world = 4: parity through the product path + bench line
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29551 tools/multi_gpu_check.py --exchange p2p > $O/r2_g4_check.log 2>&1; grep multi_gpu_check $O/r2_g4_check.log | tail -10
timeout 300 $TR --master-port 29552 tools/multi_gpu_check.py --exchange p2p --n 5000 --m 1500 --chunk 512 > $O/r2_g4_check_n5000.log 2>&1; grep multi_gpu_check $O/r2_g4_check_n5000.log | tail -3
timeout 900 $TR --master-port 29553 bench.py --gpus 4 --steps 5 --warmup 3 > $O/r2_bench_n4.json 2> $O/r2_bench_n4.err; echo "bench n4 rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_n4.json').read().strip().splitlines()[-1])
e=d["e2e"]
print("n4 ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "e2e", e and (round(e["ms_per_step"],1), e["d2h_bytes_per_step"]), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"], {k:round(v,1) for k,v in d["kernel_ms_per_step"].items()})
PY
