#!/bin/bash
# ncu --set full of the analysis kernel at n_fft 2048: the warp kernel (default) and the block kernel (SGX_K1W2=0)
mkdir -p gpurun_out
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
SGX_K1_NFR=${NFR:-48} ncu --set full --clock-control none --import-source on -k regex:stft_warp2 -s 1 -c 1 -f -o gpurun_out/prof_k1_w2 $CMD > gpurun_out/ncu_k1_w2.log 2>&1
echo "ncu w2 exit $?"
SGX_K1W2=0 ncu --set full --clock-control none --import-source on -k regex:stft_db -s 1 -c 1 -f -o gpurun_out/prof_k1_blk $CMD > gpurun_out/ncu_k1_blk.log 2>&1
echo "ncu block exit $?"
