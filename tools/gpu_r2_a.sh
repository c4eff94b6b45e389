#!/bin/bash
# round 2, step A: packed-FP32 K1 -- tests, quick benches, and the (16,2,8) CTA shape
bash tools/gpu_quick.sh
for var in "16,2,8"; do
  SGX_K1_VARIANT=$var timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_c5_v.log 2> gpurun_out/q_c5_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_c5_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("c5 variant $var step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % (d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("variant $var failed", ex); print(open("gpurun_out/q_c5_v.err").read()[-600:])
PY
done
