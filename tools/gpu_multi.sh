#!/bin/bash
# N-GPU run (NGPU, default 2): sharded parity check, NCCL path of the bench, CPU arm (SKIP_REF=1 to skip)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
N=${NGPU:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi check exit $?" | tee -a gpurun_out/multi_check.log
tail -5 gpurun_out/multi_check.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
tail -c 1500 gpurun_out/bench_n$N.log; tail -3 gpurun_out/bench_n$N.err
[ -n "$SKIP_REF" ] || timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref exit $?"; tail -c 600 gpurun_out/bench_ref_n$N.log
