#!/bin/bash
# 8-GPU session D: launch timeline at 8 ranks (limiter evidence) + C3 variants on one box
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
DGRM_TRACE=1 timeout 600 $TR --master-port 29581 bench.py --gpus 8 --steps 3 --warmup 2 --no-e2e --no-check > $O/r2_trace_n8.json 2> $O/r2_trace_n8.err; echo "trace rc $?"
python tools/trace_summary.py $O/r2_trace_n8.err > $O/r2_trace_n8_summary.txt; cat $O/r2_trace_n8_summary.txt
grep "\[dgrm trace r0\]" $O/r2_trace_n8.err | tail -60 > $O/r2_trace_n8_rank0_last_build.txt
timeout 900 $TR --master-port 29582 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e --no-check > $O/r2_bench_c3_n8_base.json 2> $O/r2_bench_c3_n8_base.err; echo "c3 base rc $?"
timeout 900 $TR --master-port 29583 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e --no-check --resident-mb 65536 > $O/r2_bench_c3_n8_res64g.json 2> $O/r2_bench_c3_n8_res64g.err; echo "c3 res64g rc $?"
for f in $O/r2_bench_c3_n8_base.json $O/r2_bench_c3_n8_res64g.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "xg", d["config"]["exchange_chunks"], d["per_rank"]["ms_per_step"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
