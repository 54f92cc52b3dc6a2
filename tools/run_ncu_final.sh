#!/bin/bash
# launch list of the bench command + ncu --set full of the final K1 kernels
cd "$(dirname "$0")/.."
O=gpurun_out
CMD="python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-check --no-k1"
$CMD > $O/r2_launches_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2_launches.csv $CMD > $O/r2_launches_ncu.log 2>&1
tail -2 $O/r2_launches_ncu.log
CMD2="python tools/gpu_perf.py --n 20000 --m 32768 --chunk 16384 --count fp4 --steps 2"
DGRM_OVERLAP=0 $CMD2 > $O/r2_k1b_plain.log 2>&1 && DGRM_OVERLAP=0 ncu --set full --clock-control none --import-source on -k regex:"bed_expand_t|bed_decode_t|diag_reduce|bed_stats_finish" -s 4 -c 4 -o $O/r2_k1b_prof -f $CMD2 > $O/r2_k1b_ncu.log 2>&1
tail -2 $O/r2_k1b_ncu.log
