#!/bin/bash
# quick check of a K1 change: parity subset, then the C5 bench under a few tuning knobs
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 -k "stft or mel or parity or golden or multitrack or ragged or slice or determinism" > gpurun_out/pytest_k1.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_k1.log
tail -4 gpurun_out/pytest_k1.log
run() {
  name=$1; shift
  env "$@" timeout 600 python bench.py --workload c5 --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/p_$name.log 2> gpurun_out/p_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/p_$name.log").read().strip().splitlines()[-1]); r=d["roofline_step"]
    print("%-22s step %.3f ms  k1 %.3f ms  k3 %.3f ms" % ("$name", d["ms_per_step"], r["k1_ms"], r["k3_ms"]))
except Exception as ex:
    print("$name failed", ex); print(open("gpurun_out/p_$name.err").read()[-800:])
PY
}
run default X=1
run nobank SGX_K1_NOBANK=1
run g4 SGX_K1_VARIANT=8,4,4
run g4_nfr16 SGX_K1_VARIANT=8,4,4 SGX_K1_NFR=16
run nobank_nfr8 SGX_K1_NOBANK=1 SGX_K1_NFR=8
