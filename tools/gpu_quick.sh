#!/bin/bash
# quick check of a kernel change: all GPU tests, then the C5 and C3 benches without the e2e / CPU legs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -5
for wl in c5 c3 c2 c1; do
  timeout 300 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_$wl.log 2> gpurun_out/q_$wl.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_$wl.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-6s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$wl", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$wl failed", ex); print(open("gpurun_out/q_$wl.err").read()[-600:])
PY
done
for F in 512 2048 8192; do
  timeout 300 python bench.py --workload c4 --n-fft $F --tracks 4 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_c4_$F.log 2> gpurun_out/q_c4_$F.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_c4_$F.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("c4_%-6s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$F", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("c4 $F failed", ex); print(open("gpurun_out/q_c4_$F.err").read()[-600:])
PY
done
