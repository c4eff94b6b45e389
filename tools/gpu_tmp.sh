#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -rA -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit" gpurun_out/pytest_gpu.log | tail -3
grep -E "wide render" gpurun_out/pytest_gpu.log | head
run() {
  name=$1; shift
  env "$@" > gpurun_out/s_$name.log 2> gpurun_out/s_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s_$name.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-22s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$name", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$name failed", ex); print(open("gpurun_out/s_$name.err").read()[-600:])
PY
}
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
run c5 X=1 $B --workload c5
run c3 X=1 $B --workload c3
