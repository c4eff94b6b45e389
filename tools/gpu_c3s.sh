#!/bin/bash
mkdir -p gpurun_out
N=${NGPU:-2}
if [ "$N" = "1" ]; then
  python bench.py --workload c3s --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c3s_n1.log 2> gpurun_out/bench_c3s_n1.err
else
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --workload c3s --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_c3s_n$N.log 2> gpurun_out/bench_c3s_n$N.err
fi
echo "c3s n$N exit $?"; tail -2 gpurun_out/bench_c3s_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_c3s_n$N.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
print("c3s N=$N value %.0f audio-s/s step %.3f ms k1 %.3f k3 %.3f scaling %s range %s"%(d["value"], d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["scaling"], d["db_range"]))
PY
