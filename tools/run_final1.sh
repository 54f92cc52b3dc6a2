#!/bin/bash
# 1-GPU final session: full GPU suite, default bench line + reference arm, C4 / C5 lines, fp16x2 lines, LD line, syrk ncu
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > $O/r2_pytest_final.log 2>&1; tail -3 $O/r2_pytest_final.log
timeout 900 python bench.py > $O/r2_bench_final_n1.json 2> $O/r2_bench_final_n1.err; echo "bench final rc $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $O/r2_bench_reference_arm.json 2> $O/r2_bench_reference_arm.err; echo "reference arm rc $?"
timeout 900 python bench.py --config c5 --steps 3 --warmup 2 --no-cpu --no-k1 --no-e2e > $O/r2_bench_c5_n1.json 2> $O/r2_bench_c5_n1.err; echo "c5 rc $?"
timeout 900 python bench.py --config c4 --steps 2 --warmup 1 --no-cpu --no-k1 > $O/r2_bench_c4_n1.json 2> $O/r2_bench_c4_n1.err; echo "c4 rc $?"
timeout 600 python bench.py --operand fp16x2 --steps 5 --warmup 3 --no-cpu --no-k1 --no-e2e > $O/r2_bench_c2_fp16x2.json 2> $O/r2_bench_c2_fp16x2.err; echo "c2 fp16x2 rc $?"
timeout 600 python bench.py --n 2000 --m 10000 --operand fp16x2 --steps 20 --warmup 3 --no-cpu --no-k1 --no-e2e > $O/r2_bench_c1_fp16x2.json 2> $O/r2_bench_c1_fp16x2.err; echo "c1 fp16x2 rc $?"
timeout 600 python bench.py --n 2000 --m 10000 --steps 20 --warmup 3 --no-cpu --no-k1 --no-e2e > $O/r2_bench_c1_fp16.json 2> $O/r2_bench_c1_fp16.err; echo "c1 fp16 rc $?"
timeout 600 python tools/gpu_ld_perf.py > $O/r2_ld_perf.json 2> $O/r2_ld_perf.err; echo "ld rc $?"; cat $O/r2_ld_perf.json | cut -c1-400
CMD="python tools/gpu_perf.py --n 20000 --m 32768 --chunk 16384 --count fp4 --steps 2"
$CMD > $O/r2_syrk_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:grm_syrk -s 3 -c 3 -o $O/r2_syrk_prof -f $CMD > $O/r2_syrk_ncu.log 2>&1; tail -2 $O/r2_syrk_ncu.log
for f in $O/r2_bench_final_n1.json $O/r2_bench_reference_arm.json $O/r2_bench_c5_n1.json $O/r2_bench_c4_n1.json $O/r2_bench_c2_fp16x2.json $O/r2_bench_c1_fp16x2.json $O/r2_bench_c1_fp16.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    e=d.get("e2e")
    print(sys.argv[1].split('/')[-1], "ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", e and round(e.get("ms_per_step",0),1), "parity", d.get("parity") and d["parity"]["ok"], d.get("parity") and d["parity"]["A_max_err"], "checks", d.get("fullsize_checks") and d["fullsize_checks"]["ok"], "cpu", d.get("cpu_baseline") and d["cpu_baseline"].get("value"))
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
