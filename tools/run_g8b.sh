#!/bin/bash
# 8-GPU session B: e2e transports at 8 ranks, C3 exchange granularity on one box
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
for A in fp64 fp32; do
DGRM_TRACE_OUT=1 DGRM_BUILD_TRACE=1 timeout 600 $TR --master-port 29541 bench.py --gpus 8 --steps 3 --warmup 3 --no-check --a-transport $A --n-transport narrow > $O/r2_bench_n8_${A}_narrow.json 2> $O/r2_bench_n8_${A}_narrow.err; echo "$A narrow rc $?"; grep "\[multi r0\]\|\[dgrm\]" $O/r2_bench_n8_${A}_narrow.err | tail -2
done
timeout 900 $TR --master-port 29542 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e --no-check --xg 4 > $O/r2_bench_c3_n8_p2p_xg4.json 2> $O/r2_bench_c3_n8_p2p_xg4.err; echo "c3 p2p xg4 rc $?"
timeout 900 $TR --master-port 29543 bench.py --gpus 8 --config c3 --steps 3 --warmup 2 --no-e2e --no-check --exchange nccl > $O/r2_bench_c3_n8_nccl_xg4.json 2> $O/r2_bench_c3_n8_nccl_xg4.err; echo "c3 nccl xg4 rc $?"
for f in $O/r2_bench_n8_*_narrow.json $O/r2_bench_c3_n8_*xg4.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d["e2e"]
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e pct_burst %.1f"%(d["ms_per_step"], d["value"], d["pct_of_bf16_peak"]["burst"]), "e2e", e and (round(e["ms_per_step"],1), e["d2h_bytes_per_step"], e.get("roofline_e2e",{}).get("frac")), "xg", d["config"]["exchange_chunks"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
