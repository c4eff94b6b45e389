#!/bin/bash
# ncu launch list + full captures of K1 / K3 on a reduced C5 batch (same kernels, 4 tracks)
mkdir -p gpurun_out
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -c 600 gpurun_out/plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "ncu launches exit $?"
ncu --set full --clock-control none --import-source on -k regex:stft_db -s 1 -c 1 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu k1 exit $?"
ncu --set full --clock-control none --import-source on -k regex:render_fast -s 1 -c 1 -f -o gpurun_out/prof_k3 $CMD > gpurun_out/ncu_k3.log 2>&1
echo "ncu k3 exit $?"
ls -la gpurun_out
