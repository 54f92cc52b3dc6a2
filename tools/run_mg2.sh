#!/bin/bash
# 2-GPU validation session (gpurun --gpus 2): parity through the product's exchange loop, then bench variants.
cd "$(dirname "$0")/.."
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_gpu_ld.py tests/test_gpu_multi.py -q > $O/r2_mg2_pytest.log 2>&1; tail -3 $O/r2_mg2_pytest.log
timeout 300 $TR --master-port 29511 tools/multi_gpu_check.py --exchange nccl > $O/r2_mg2_check_nccl.log 2>&1; grep multi_gpu_check $O/r2_mg2_check_nccl.log | tail -12
DGRM_EXCHANGE_MAP=ipc timeout 300 $TR --master-port 29512 tools/multi_gpu_check.py --exchange p2p > $O/r2_mg2_check_p2p_ipc.log 2>&1; grep "multi_gpu_check.*ALL\|FAIL" $O/r2_mg2_check_p2p_ipc.log | tail -3
DGRM_EXCHANGE_MAP=symm timeout 300 $TR --master-port 29513 tools/multi_gpu_check.py --exchange p2p > $O/r2_mg2_check_p2p_symm.log 2>&1; grep "multi_gpu_check.*ALL\|FAIL\|Error" $O/r2_mg2_check_p2p_symm.log | tail -3
timeout 300 $TR --master-port 29514 tools/d2h_concurrency_probe.py --mb 512 --out $O/r2_d2h_concurrency_n2.json > $O/r2_mg2_probe.log 2>&1; grep "ranks" $O/r2_mg2_probe.log
timeout 600 $TR --master-port 29515 bench.py --gpus 2 --steps 5 --warmup 3 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err; echo "bench n2 rc $?"
timeout 600 $TR --master-port 29516 bench.py --gpus 2 --steps 5 --warmup 3 --xg 1 --no-e2e --no-check > $O/r2_bench_n2_xg1.json 2> $O/r2_bench_n2_xg1.err; echo "bench n2 xg1 rc $?"
DGRM_EXCHANGE_MAP=ipc timeout 600 $TR --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --exchange p2p --no-e2e --no-check > $O/r2_bench_n2_p2p_ipc.json 2> $O/r2_bench_n2_p2p_ipc.err; echo "bench n2 p2p ipc rc $?"
DGRM_EXCHANGE_MAP=symm timeout 600 $TR --master-port 29518 bench.py --gpus 2 --steps 5 --warmup 3 --exchange p2p --no-e2e --no-check > $O/r2_bench_n2_p2p_symm.json 2> $O/r2_bench_n2_p2p_symm.err; echo "bench n2 p2p symm rc $?"
timeout 600 $TR --master-port 29519 bench.py --gpus 2 --steps 5 --warmup 3 --a-transport fp64 --no-check > $O/r2_bench_n2_fp64.json 2> $O/r2_bench_n2_fp64.err; echo "bench n2 fp64 rc $?"
for f in $O/r2_bench_n2*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"], "xg", d["config"]["exchange_chunks"])
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
