#!/bin/bash
# 8-GPU session: parity through the product path, the weak-scaling bench line with its parity block, C3 at full
# size, the host-link probe, and a 50k GRM written by 8 engines through the host CLI.
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
nproc > $O/r2_g8_host.txt; free -g >> $O/r2_g8_host.txt; nvidia-smi topo -m >> $O/r2_g8_host.txt 2>&1
timeout 300 $TR --master-port 29531 tools/multi_gpu_check.py > $O/r2_g8_check.log 2>&1; grep multi_gpu_check $O/r2_g8_check.log | tail -12
timeout 300 $TR --master-port 29532 tools/d2h_concurrency_probe.py --mb 512 --out $O/r2_d2h_concurrency_n8.json > $O/r2_g8_probe.log 2>&1; grep "ranks" $O/r2_g8_probe.log
DGRM_TRACE_OUT=1 DGRM_BUILD_TRACE=1 timeout 900 $TR --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > $O/r2_bench_n8.json 2> $O/r2_bench_n8.err; echo "bench n8 rc $?"; grep "\[multi r0\]\|\[dgrm\]" $O/r2_bench_n8.err | tail -2
timeout 900 $TR --master-port 29534 bench.py --gpus 8 --steps 5 --warmup 3 --exchange nccl --no-check --no-e2e > $O/r2_bench_n8_nccl.json 2> $O/r2_bench_n8_nccl.err; echo "bench n8 nccl rc $?"
timeout 900 $TR --master-port 29535 bench.py --gpus 8 --steps 3 --warmup 2 --a-transport fp64 --n-transport fp64 --no-check > $O/r2_bench_n8_direct.json 2> $O/r2_bench_n8_direct.err; echo "bench n8 direct rc $?"
timeout 1200 $TR --master-port 29536 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e > $O/r2_bench_c3_n8.json 2> $O/r2_bench_c3_n8.err; echo "bench c3 rc $?"
timeout 900 python tools/cli_bigfile_check.py --n 50000 --m 20000 --engines 8 --dir /tmp/c5cli8 > $O/r2_cli_50k_8engines.json 2> $O/r2_cli_50k_8engines.err; echo "cli rc $?"; cat $O/r2_cli_50k_8engines.json
for f in $O/r2_bench_n8*.json $O/r2_bench_c3_n8.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d["e2e"]
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "e2e", e and (round(e["ms_per_step"],1), e["d2h_bytes_per_step"]), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
