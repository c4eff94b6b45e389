#!/bin/bash
# round 2, step B: warp-per-frame-pair K1 (W2) -- tests, quick benches, W2 vs block kernel, tile sizes
bash tools/gpu_quick.sh
run() { # label, env...
  label=$1; shift
  env "$@" timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/q_c5_v.log 2> gpurun_out/q_c5_v.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/q_c5_v.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("c5 $label step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % (d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$label failed", ex); print(open("gpurun_out/q_c5_v.err").read()[-600:])
PY
}
run "block kernel" SGX_K1W2=0
run "w2 16 frames/tile" SGX_K1_NFR=16
run "w2 48 frames/tile" SGX_K1_NFR=48
run "w2 64 frames/tile" SGX_K1_NFR=64
