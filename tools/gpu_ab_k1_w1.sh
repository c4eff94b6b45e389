#!/bin/bash
# round 2: the warp-per-frame kernel on packed complex values (SGX_K1W1=1) -- its parity test, then C5 with 16 / 12 / 8 warps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_properties.py -m gpu -q --timeout 600 -k "n_fft_2048" > gpurun_out/pytest_w1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_w1.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_w1.log | tail -8
run() { # label, env, workload args
  label=$1; shift; e=$1; shift
  env $e timeout 300 python bench.py "$@" --steps 5 --warmup 3 --no-cpu --no-configs --no-e2e > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-22s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run "c5 W1 16 warps" "SGX_K1W1=1" --workload c5
run "c5 W1 12 warps" "SGX_K1W1=1 SGX_W1_WARPS=12" --workload c5
run "c5 block" "SGX_K1W1=0" --workload c5
