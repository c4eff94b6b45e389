#!/bin/bash
# K1 per FFT size (C4 points, 4 tracks) under the default and an alternative CTA shape; plus C3.
# The alternative shapes exist only in a library built with `make TUNE=-DSGX_K1_ALTERNATES`.
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
B="timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e"
for F in 512 1024 2048 4096 8192 16384; do
  run c4_${F} X=1 $B --workload c4 --n-fft $F --tracks 4
done
run c4_512_g8 SGX_K1_VARIANT=8,4,8 $B --workload c4 --n-fft 512 --tracks 4
run c4_1024_g8 SGX_K1_VARIANT=8,4,8 $B --workload c4 --n-fft 1024 --tracks 4
run c4_1024_g4 SGX_K1_VARIANT=8,4,4 $B --workload c4 --n-fft 1024 --tracks 4
run c4_2048_g2 SGX_K1_VARIANT=8,4,2 $B --workload c4 --n-fft 2048 --tracks 4
run c4_4096_g2 SGX_K1_VARIANT=8,4,2 $B --workload c4 --n-fft 4096 --tracks 4
run c4_4096_g1 SGX_K1_VARIANT=8,4,1 $B --workload c4 --n-fft 4096 --tracks 4
run c3 X=1 $B --workload c3
run c3_g2 SGX_K1_VARIANT=8,4,2 $B --workload c3
run c3_g1 SGX_K1_VARIANT=8,4,1 $B --workload c3
