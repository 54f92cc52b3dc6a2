#!/bin/bash
# final single-GPU verification of the round: whole GPU suite, smoke, the default bench line, dosage / BGEN perf
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -6
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 200 python tools/gpu_dosage_perf.py > gpurun_out/r2_dosage_perf.json 2> gpurun_out/r2_dosage_perf.err; echo perf rc $?
cat gpurun_out/r2_dosage_perf.json
timeout 400 python bench.py > gpurun_out/r2_bench_final3_n1.json 2> gpurun_out/r2_bench_final3_n1.err; echo bench rc $?
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_final3_n1.json").read().strip().splitlines()[-1])
print("bench ms", d["ms_per_step"], "value %.3e" % d["value"], "e2e", d.get("e2e", {}).get("value"), "parity", d.get("parity", {}).get("ok"),
      "clocks", d.get("clocks"))
PY
