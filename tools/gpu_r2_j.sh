#!/bin/bash
# round 2, step J: render launches grouped by tiling plan + range commit fused into the reduce launch --
# full GPU suite, then quick device-only bench lines of every config (launch counts included)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|AssertionError:" gpurun_out/pytest_gpu.log | tail -12
run() { # label, workload args
  label=$1; shift
  timeout 300 python bench.py "$@" --steps 5 --warmup 3 --no-cpu --no-configs --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-12s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f  launches/step %.1f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"], d["gpu_launches"]/d["steps"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run c5 --workload c5
run c3 --workload c3
run c4_512 --workload c4 --n-fft 512 --tracks 4
run c4_16384 --workload c4 --n-fft 16384 --tracks 4
run c2 --workload c2
run c1 --workload c1
