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
    print("variant %-8s nfr %-3s  k1 %.3f ms  k3 %.3f ms  step %.3f ms range %s" % (sys.argv[1], sys.argv[2], r["k1_ms"], r["k3_ms"], d["ms_per_step"], d["db_range"]))
except Exception as e:
    print("variant", sys.argv[1], sys.argv[2], "FAILED", e); print(open("gpurun_out/sweep.err").read()[-800:])
PY
}
for v in ${VARIANTS:-"8,4,2" "16,2,4" "16,2,2"}; do
  for nfr in ${NFRS:-0}; do run "$v" "$nfr"; done
done
