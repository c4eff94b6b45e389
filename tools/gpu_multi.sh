#!/bin/bash
# N-GPU run (NGPU, default 2): sharded parity check (NCCL inside libsgx.so), the bench over N ranks, the C++ bench with ONE
# handle over all GPUs (no Python), the CPU arm (SKIP_REF=1 to skip)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/multi_gpus.txt
N=${NGPU:-2}
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/multi_gpu_check.py > gpurun_out/multi_check.log 2>&1; echo "multi check exit $?" | tee -a gpurun_out/multi_check.log
grep -E "identical|range|PASSED|FAILED|Error|error" gpurun_out/multi_check.log | tail -24
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.log 2> gpurun_out/bench_n$N.err; echo "bench n$N exit $?"
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_n$N.log").read().strip().splitlines()[-1]); r=d["roofline_step"]; e=d["e2e"]
    print("n=$N value %.0f step %.3f ms k1 %.3f k3 %.3f launches %d" % (d["value"], d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["gpu_launches"]))
    print("e2e %.0f (%.1f ms)  one batch at a time %.0f (%.1f ms)  int16 pipelined %.0f" % (e["value"], e["ms_per_step"], e["one_batch_at_a_time"]["value"], e["one_batch_at_a_time"]["ms_per_step"], e["int16_pcm_pipelined"]["value"]))
    print("link", e["per_rank_link_gbs"], e["one_batch_at_a_time"]["per_rank_link_gbs"])
except Exception as ex:
    print("bench parse failed", ex); print(open("gpurun_out/bench_n$N.err").read()[-1500:])
PY
timeout 900 ./benches/bench --c5 $((32 * N)) 600 2>&1 | tail -2
if [ -z "$SKIP_REF" ]; then timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.log 2>&1; echo "ref exit $?"; tail -c 400 gpurun_out/bench_ref_n$N.log; fi
