#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2_pytest_verify.log 2>&1; tail -3 $O/r2_pytest_verify.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > $O/r2_bench_verify_n1.json 2> $O/r2_bench_verify_n1.err; echo "bench rc $?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_verify_n1.json').read().strip().splitlines()[-1])
print("ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", round(d["e2e"]["ms_per_step"],1), "parity", d["parity"]["ok"], "checks", d["fullsize_checks"]["ok"], "k1", {k:(round(v["ms_per_step"],3), round(v["frac"],3)) for k,v in d["roofline_k1"]["kernels"].items()}, "cpu %.3e"%d["cpu_baseline"]["value"], "launches", d["gpu_launches"])
PY
