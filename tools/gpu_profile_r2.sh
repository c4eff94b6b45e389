#!/bin/bash
# round-2 ncu evidence on a reduced C5 batch (same kernels, 4 tracks), each pass only after the plain run exited 0:
# launch list, --set full of K1 (stft_db_kernel) and of K3 (render_slide_kernel)
mkdir -p gpurun_out
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -c 300 gpurun_out/plain.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:stft_db -s 1 -c 1 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1; echo "ncu k1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:render_slide -s 1 -c 1 -f -o gpurun_out/prof_k3 $CMD > gpurun_out/ncu_k3.log 2>&1; echo "ncu k3 exit $?"
