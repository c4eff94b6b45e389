#!/bin/bash
# ncu --set full of the sliding-window render kernel at the bench geometry (4 tracks), after a plain run
mkdir -p gpurun_out
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu --no-configs"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:render_slide -s 1 -c 1 -f -o gpurun_out/prof_k3_slide $CMD > gpurun_out/ncu_k3_slide.log 2>&1
echo "ncu exit $?"
