#!/bin/bash
# 1-GPU session: LD geometry reuse, K1 A/B on one box, full suite once more
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2_pytest_final2.log 2>&1; tail -3 $O/r2_pytest_final2.log
timeout 600 python tools/gpu_ld_perf.py > $O/r2_ld_perf2.json 2> $O/r2_ld_perf2.err; echo "ld rc $?"; cut -c1-300 $O/r2_ld_perf2.json
for rep in 1 2; do
DGRM_OVERLAP=0 python tools/gpu_perf.py --n 20000 --m 98304 --chunk 16384 --count fp4 --steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('transposed', d['kernel_ms'])"
DGRM_K1_TRANSPOSED=0 DGRM_OVERLAP=0 python tools/gpu_perf.py --n 20000 --m 98304 --chunk 16384 --count fp4 --steps 3 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('rowmajor  ', d['kernel_ms'])"
done
timeout 900 python bench.py --no-cpu > $O/r2_bench_final2_n1.json 2> $O/r2_bench_final2_n1.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final2_n1.json').read().strip().splitlines()[-1])
print("ms %.2f"%d["ms_per_step"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "k1", {k:(round(v["ms_per_step"],3), round(v["frac"],3)) for k,v in d["roofline_k1"]["kernels"].items()}, d["roofline_k1"]["write_only_stream_gbs"])
PY
