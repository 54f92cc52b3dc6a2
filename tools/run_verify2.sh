#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > $O/r2_pytest_verify2.log 2>&1; tail -3 $O/r2_pytest_verify2.log
timeout 300 $TR --master-port 29591 tools/multi_gpu_check.py --exchange p2p --individuals 5000 --snps 1500 --chunk 512 > $O/r2_v2_check.log 2>&1; grep multi_gpu_check $O/r2_v2_check.log | tail -3
timeout 900 $TR --master-port 29592 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2_bench_verify_n2.json 2> $O/r2_bench_verify_n2.err; echo "bench n2 rc $?"
timeout 900 $TR --master-port 29593 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > $O/r2_bench_verify_ref_n2.json 2> $O/r2_bench_verify_ref_n2.err; echo "reference arm under torchrun rc $?"; cut -c1-200 $O/r2_bench_verify_ref_n2.json
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_verify_n2.json').read().strip().splitlines()[-1])
print("n2 ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "parity", d["parity"]["ok"], "checks", d["fullsize_checks"]["ok"], d["per_rank"]["ms_per_step"])
PY
