#!/bin/bash
# round 2, step G: loader instantiations, robust truth gate, K3 variants (constant LUT / FFMA2)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -8
run() { # label, workload args, env...
  label=$1; shift; wl=$1; shift
  env "$@" timeout 300 python bench.py $wl --steps 5 --warmup 3 --no-cpu --no-configs --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-28s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run "c5 k3 var 0" "--workload c5" SGX_K3_VAR=0
run "c5 k3 var 1 (const LUT)" "--workload c5" SGX_K3_VAR=1
run "c5 k3 var 2 (ffma2)" "--workload c5" SGX_K3_VAR=2
run "c5 k3 var 3 (both)" "--workload c5" SGX_K3_VAR=3
run "c3 (stereo mel)" "--workload c3" A=1
SGX_K3_VAR=3 timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "golden or six_rates or c5_track" > gpurun_out/pytest_k3var.log 2>&1; echo "pytest(k3 var 3) exit $?"; tail -2 gpurun_out/pytest_k3var.log
