#!/bin/bash
# sliding-window render kernel (default) against render_fast_kernel (SGX_K3_SLIDE=0): GPU suite, C5 / C3 / C2 / C1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|AssertionError:|Error" gpurun_out/pytest_gpu.log | tail -12
run() { # label, env, workload args
  label=$1; shift; e=$1; shift
  env $e timeout 300 python bench.py "$@" --steps 5 --warmup 3 --no-cpu --no-configs --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-22s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f  launches/step %.1f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"], d["gpu_launches"]/d["steps"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run "c5 slide" SGX_K3_SLIDE=1 --workload c5
run "c3 slide" SGX_K3_SLIDE=1 --workload c3
run "c5 fast" SGX_K3_SLIDE=0 --workload c5
run "c2 slide" SGX_K3_SLIDE=1 --workload c2
run "c1 slide" SGX_K3_SLIDE=1 --workload c1
