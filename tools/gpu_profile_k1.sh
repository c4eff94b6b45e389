#!/bin/bash
mkdir -p gpurun_out
CMD="python bench.py --tracks 4 --steps 2 --warmup 1 --no-e2e --no-cpu ${BENCH_ARGS}"
$CMD > gpurun_out/plain.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-stft_db} -s 1 -c 1 -f -o gpurun_out/prof_${KNAME:-k1} $CMD > gpurun_out/ncu_k1.log 2>&1
echo "ncu exit $?"
