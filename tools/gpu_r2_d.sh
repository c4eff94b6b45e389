#!/bin/bash
# round 2, step D: unrolled segment mel; W2 with 8 / 10 / 12 warps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -5
run() { # label, workload args, env...
  label=$1; shift; wl=$1; shift
  env "$@" timeout 300 python bench.py $wl --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-28s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run "c5 w2 8 warps 48 fr" "--workload c5" SGX_K1_NFR=48
run "c5 w2 8 warps 32 fr" "--workload c5" SGX_K1_NFR=32
run "c5 w2 10 warps 40 fr" "--workload c5" SGX_W2_WARPS=10 SGX_K1_NFR=40
run "c5 w2 12 warps 24 fr" "--workload c5" SGX_W2_WARPS=12
run "c5 block" "--workload c5" SGX_K1W2=0
run "c3" "--workload c3" A=1
run "c2" "--workload c2" A=1
run "c1" "--workload c1" A=1
SGX_W2_WARPS=12 timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "parity or golden" > gpurun_out/pytest_gpu12.log 2>&1; echo "pytest(12 warps) exit $?"; tail -2 gpurun_out/pytest_gpu12.log
