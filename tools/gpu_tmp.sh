#!/bin/bash
mkdir -p gpurun_out
run() {
  name=$1; shift
  env "$@" > gpurun_out/s_$name.log 2> gpurun_out/s_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/s_$name.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-22s step %.3f ms  k1 %.3f ms  k3 %.3f ms  value %.0f" % ("$name", d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["value"]))
except Exception as ex:
    print("$name failed", ex); print(open("gpurun_out/s_$name.err").read()[-600:])
PY
}
B="timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e"
for fx in 4 7 10 14 20; do
run c5_fixed$fx SGX_MEL_FIXED=$fx $B --workload c5
done
run c3_fixed4 SGX_MEL_FIXED=4 $B --workload c3
run c3_fixed10 SGX_MEL_FIXED=10 $B --workload c3
run c3_fixed20 SGX_MEL_FIXED=20 $B --workload c3
