#!/bin/bash
# rebuild render_kernel.cu with different tuning macros on the GPU box and time K3 (8 tracks of C5)
cd multi-spectrogram-viewer_b200
for cfg in "4 4" "8 4" "8 3" "11 3"; do
  set -- $cfg
  rm -f build/render_kernel.cu.o
  make -s TUNE="-DSGX_K3_ABATCH=$1 -DSGX_K3_CTAS=$2" > /dev/null 2>&1 || { echo "build failed $cfg"; continue; }
  out=$(cd .. && python bench.py --tracks 8 --steps 5 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1)
  python -c "
import json
d=json.loads('''$out'''); r=d['roofline_step']; print('ABATCH=$1 CTAS=$2: k3 %.3f ms (k1 %.3f)'%(r['k3_ms'],r['k1_ms']))"
done
