#!/bin/bash
for d in 0 1 2 3 4 8 9 11 15; do
  out=$(SGX_K1_DEBUG=$d python bench.py --tracks 8 --steps 5 --warmup 2 --no-e2e --no-cpu 2>/dev/null | tail -1)
  python -c "
import json,sys
d=json.loads('''$out'''); r=d['roofline_step']; print('debug $d: k1 %.3f k3 %.3f'%(r['k1_ms'],r['k3_ms']))"
done
