#!/bin/bash
# 1-GPU session: K1 kernels alone (overlap off) + one ncu --set full capture of the decode / expand kernels
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python tools/gpu_perf.py --n 20000 --m 32768 --chunk 16384 --count fp4 --steps 2"
DGRM_OVERLAP=0 $CMD > $O/r2_k1_plain.log 2>&1 && \
DGRM_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"bed_expand_t|bed_decode_t|diag_reduce" -s 3 -c 6 -o $O/r2_k1_prof -f $CMD > $O/r2_k1_ncu.log 2>&1
tail -3 $O/r2_k1_plain.log; tail -3 $O/r2_k1_ncu.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-check > $O/r2_bench4.json 2> $O/r2_bench4.err; echo "bench4 rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench4.json').read().strip().splitlines()[-1])
print("ms %.2f"%d["ms_per_step"], "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1))
print("   k1", {k:(round(v["ms_per_step"],3), round(v["frac"],3)) for k,v in d["roofline_k1"]["kernels"].items()})
PY
