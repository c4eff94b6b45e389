#!/bin/bash
# N-GPU sharded parity check only (NGPU, default 2)
mkdir -p gpurun_out
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi check exit $?" | tee -a gpurun_out/multi_check.log
grep -E "identical|PASSED|FAILED|Error|error" gpurun_out/multi_check.log | tail -12
