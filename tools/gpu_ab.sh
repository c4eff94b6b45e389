#!/bin/bash
for i in 1 2; do
for t in 9 16; do
  out=$(SGX_MIN_TAPS=$t python bench.py --tracks 8 --steps 5 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1)
  python -c "
import json,sys
d=json.loads('''$out'''); r=d['roofline_step']; print('min_taps $t: k1 %.3f k3 %.3f'%(r['k1_ms'],r['k3_ms']))"
done; done
