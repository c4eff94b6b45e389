#!/bin/bash
# round 2, step F: raw-tile loader (int16 / linear stereo via TMA), truth-anchored dB gates, device restore: tests + quick benches
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -8
run() { # label, workload args, env...
  label=$1; shift; wl=$1; shift
  env "$@" timeout 300 python bench.py $wl --steps 5 --warmup 3 --no-cpu --no-configs > gpurun_out/q_v.log 2> gpurun_out/q_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]; e=d.get("e2e") or {}
    print("%-28s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f  e2e %s pipelined %s int16 %s" % ("$label", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"], e.get("value"), (e.get("pipelined") or {}).get("value"), (e.get("int16_pcm_pipelined") or {}).get("value")))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_v.err").read()[-600:])
PY
}
run "c5" "--workload c5" A=1
run "c3 (stereo mel, raw tiles)" "--workload c3 --no-e2e" A=1
run "c4 2048 stereo linear" "--workload c4 --n-fft 2048 --tracks 4 --channels 2 --no-e2e" A=1
run "c4 2048 mono linear" "--workload c4 --n-fft 2048 --tracks 4 --no-e2e" A=1
