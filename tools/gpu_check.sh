#!/bin/bash
# parity tests + benches (c5 headline, c3, c2, c1 and the six points of the c4 FFT-size sweep)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -q -rA -s --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|error" gpurun_out/pytest_gpu.log | tail -5
grep -E "^FAILED" gpurun_out/pytest_gpu.log | head -20
for wl in c5 c3 c2 c1 c4_512 c4_1024 c4_2048 c4_4096 c4_8192 c4_16384; do
  extra=""; [ "$wl" != "c5" ] && extra="--no-cpu --no-e2e"
  [ -n "$BENCH_FAST" ] && extra="--no-cpu --no-e2e"
  sel="--workload $wl"
  case $wl in c4_*) sel="--workload c4 --n-fft ${wl#c4_} --tracks 4";; esac
  timeout 900 python bench.py $sel --steps 5 --warmup 3 $extra > gpurun_out/bench_$wl.log 2> gpurun_out/bench_$wl.err; echo "bench $wl exit $?" >> gpurun_out/bench_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_$wl.log").read().strip().splitlines()[-1])
    r=d["roofline_step"]; e=d.get("e2e") or {}; c=d.get("cpu_baseline") or {}
    print("$wl value %.0f audio-s/s  step %.3f ms  k1 %.3f ms  k3 %.3f ms  step-frac %.3f  k1-frac %.3f  e2e %s  cpu %s launches %s" % (d["value"], d["ms_per_step"], r["k1_ms"], r["k3_ms"], r["frac"], d["roofline"]["frac"], e.get("value"), c.get("value"), d["gpu_launches"]))
except Exception as ex:
    print("$wl bench parse failed", ex); print(open("gpurun_out/bench_$wl.err").read()[-1500:])
PY
done
