#!/bin/bash
# 1-GPU session: full GPU test suite, then K1 A/B (transposed decode + pair-table expand vs round-1 path)
cd "$(dirname "$0")/.."
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q > $O/r2_pytest3.log 2>&1; tail -4 $O/r2_pytest3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r2_bench3.json 2> $O/r2_bench3.err; echo "bench3 rc $?"
DGRM_K1_TRANSPOSED=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu --no-check --no-e2e > $O/r2_bench3_rowmajor.json 2> $O/r2_bench3_rowmajor.err; echo "rowmajor rc $?"
for f in $O/r2_bench3*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms %.2f value %.3e"%(d["ms_per_step"], d["value"]), "e2e", d["e2e"] and round(d["e2e"]["ms_per_step"],1), "parity", d["parity"] and d["parity"]["ok"], "checks", d["fullsize_checks"] and d["fullsize_checks"]["ok"])
    print("   k1", {k:(round(v["ms_per_step"],3), round(v["frac"],3)) for k,v in d["roofline_k1"]["kernels"].items()})
except Exception as ex:
    print(sys.argv[1], "unreadable:", ex)
PY
done
