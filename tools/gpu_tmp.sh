#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 -k "render or geometry or image or wav or slice or golden or grey" > gpurun_out/pytest_k3.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k3.log
tail -3 gpurun_out/pytest_k3.log
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
B="timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e"
for F in 512 4096 8192 16384; do
  run c4_${F} X=1 $B --workload c4 --n-fft $F --tracks 4
  run c4_${F}_py32 SGX_K3_WIDE_PY=32 $B --workload c4 --n-fft $F --tracks 4
  run c4_${F}_py16 SGX_K3_WIDE_PY=16 $B --workload c4 --n-fft $F --tracks 4
done
for F in 4096 8192 16384; do
  run c4_${F}_px128 SGX_K3_WIDE_PX=128 $B --workload c4 --n-fft $F --tracks 4
  run c4_${F}_px256 SGX_K3_WIDE_PX=256 $B --workload c4 --n-fft $F --tracks 4
  run c4_${F}_px512 SGX_K3_WIDE_PX=512 $B --workload c4 --n-fft $F --tracks 4
done
run c4_512_px128 SGX_K3_WIDE_PX=128 $B --workload c4 --n-fft 512 --tracks 4
