#!/bin/bash
# round 2, step H: table twiddles (accuracy) -- tests incl. the f64-truth gates, quick benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|AssertionError:" gpurun_out/pytest_gpu.log | tail -12
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
run "c5" "--workload c5" SGX_K3_VAR=1
run "c3" "--workload c3" A=1
run "c4 512" "--workload c4 --n-fft 512 --tracks 4" A=1
run "c4 16384" "--workload c4 --n-fft 16384 --tracks 4" A=1
