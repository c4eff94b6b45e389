#!/bin/bash
# round 2, step I: tcgen05 render path -- parity tests, C5 / C2 benches with and without it
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|AssertionError:|Error" gpurun_out/pytest_gpu.log | tail -12
run() { # label, workload args, env...
  label=$1; shift; wl=$1; shift
  env "$@" timeout 300 python bench.py $wl --steps 5 --warmup 3 --no-cpu --no-configs --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-28s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-900:])
PY
}
run "c5 tc" "--workload c5" A=1
run "c5 fp32" "--workload c5" SGX_K3_TC=0
run "c2 tc" "--workload c2" A=1
run "c2 fp32" "--workload c2" SGX_K3_TC=0
