#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
P="python tools/gpu_perf.py --n 20000 --m 98304 --chunk 16384 --count fp4 --steps 3"
show() { tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', {k: d['kernel_ms'][k] for k in ('expand','gather','syrk','paircnt')}, round(d['wall_ms'],1))"; }
for rep in 1 2; do
DGRM_OVERLAP=0 $P 2>&1 | show "base      "
DGRM_EXPAND_KFAST=1 DGRM_OVERLAP=0 $P 2>&1 | show "kfast     "
DGRM_EXPAND_CS=1 DGRM_OVERLAP=0 $P 2>&1 | show "cs        "
DGRM_EXPAND_KFAST=1 DGRM_EXPAND_CS=1 DGRM_OVERLAP=0 $P 2>&1 | show "kfast+cs  "
done
$P 2>&1 | show "overlap base    "
DGRM_EXPAND_KFAST=1 $P 2>&1 | show "overlap kfast   "
DGRM_EXPAND_CS=1 $P 2>&1 | show "overlap cs      "
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-check --no-k1"
$CMD > $O/r2_launches_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"grm_syrk|bed_|diag_reduce|fix_diag|pair_count" -c 600 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_launches_ncu.log 2>&1
tail -2 $O/r2_launches_ncu.log
