#!/bin/bash
# ncu evidence on a reduced C5 batch (same kernels, 4 tracks).  ONE ncu pass per call:
#   STEP=launches  launch list (gpu__time_duration per launch)
#   STEP=k1 | k3   --set full capture of the analysis / render kernel (-> gpurun_out/prof_<step>.ncu-rep)
mkdir -p gpurun_out
STEP=${STEP:-launches}
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
tail -c 400 gpurun_out/plain.log
case $STEP in
  launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1;;
  k1) ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-stft_db} -s 1 -c 1 -f -o gpurun_out/prof_k1 $CMD > gpurun_out/ncu_k1.log 2>&1;;
  k3) ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-render_fast} -s 1 -c 1 -f -o gpurun_out/prof_k3 $CMD > gpurun_out/ncu_k3.log 2>&1;;
esac
echo "ncu $STEP exit $?"
