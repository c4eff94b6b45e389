#!/bin/bash
mkdir -p gpurun_out
export SGX_K1W1=1
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:stft_warp1 -s 1 -c 1 -f -o gpurun_out/prof_k1_w1 $CMD > gpurun_out/ncu_k1_w1.log 2>&1; echo "ncu exit $?"
