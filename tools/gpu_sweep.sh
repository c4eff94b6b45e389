#!/bin/bash
# K1 variant sweep on a reduced C5 batch (8 tracks): prints k1 / k3 ms per variant
mkdir -p gpurun_out
run() {
  out=$(SGX_K1_VARIANT="$1" SGX_K1_NFR="$2" timeout 300 python bench.py --tracks 8 --steps 4 --warmup 2 --no-e2e --no-cpu 2>gpurun_out/sweep.err | tail -1)
  python - "$1" "$2" <<PY
import json,sys
try:
    d=json.loads('''$out''')
    r=d["roofline_step"]
    print("variant %-8s nfr %-3s  k1 %.3f ms  k3 %.3f ms  step %.3f ms" % (sys.argv[1], sys.argv[2], r["k1_ms"], r["k3_ms"], d["ms_per_step"]))
except Exception as e:
    print("variant", sys.argv[1], sys.argv[2], "FAILED", e); print(open("gpurun_out/sweep.err").read()[-800:])
PY
}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 600 2>&1 | tail -2
for v in "8,4,2" "8,2,2" "8,2,4" "4,4,1" "4,4,2"; do
  for nfr in 0 8 4; do run "$v" "$nfr"; done
done
