#!/bin/bash
# 2-GPU session B: e2e anatomy (transports, traces), default bench line, 2-rank pytest
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 900 python -m pytest tests/test_gpu_multi.py -q > $O/r2_mg2b_pytest.log 2>&1; tail -3 $O/r2_mg2b_pytest.log
timeout 600 $TR --master-port 29521 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2b_bench_n2.json 2> $O/r2b_bench_n2.err; echo "default rc $?"
for A in fp64 fp32; do for N in narrow fp64; do
DGRM_TRACE_OUT=1 DGRM_BUILD_TRACE=1 timeout 600 $TR --master-port 29522 bench.py --gpus 2 --steps 4 --warmup 3 --no-check --a-transport $A --n-transport $N > $O/r2b_bench_n2_${A}_${N}.json 2> $O/r2b_bench_n2_${A}_${N}.err; echo "$A $N rc $?"; grep "\[multi r0\]\|\[dgrm\]" $O/r2b_bench_n2_${A}_${N}.err | tail -3
done; done
timeout 600 $TR --master-port 29523 bench.py --gpus 2 --steps 5 --warmup 3 --chunk 32768 --no-check --no-e2e > $O/r2b_bench_n2_chunk32k.json 2> $O/r2b_bench_n2_chunk32k.err; echo "chunk32k rc $?"
timeout 600 $TR --master-port 29524 bench.py --gpus 2 --steps 5 --warmup 3 --exchange nccl --no-check --no-e2e > $O/r2b_bench_n2_nccl.json 2> $O/r2b_bench_n2_nccl.err; echo "nccl rc $?"
for f in $O/r2b_bench_n2*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d["e2e"]
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", e and (round(e["ms_per_step"],1), e["d2h_bytes_per_step"]), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"], d["config"]["parallelism"][:60])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
