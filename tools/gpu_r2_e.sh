#!/bin/bash
# round 2, step E: C-ABI multi-GPU plumbing + batched host images on ONE GPU: tests, full default bench, C++ bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|pytest exit|Error|assert" gpurun_out/pytest_gpu.log | tail -5
timeout 900 python bench.py > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench exit $?"
python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/bench_default.log").read().strip().splitlines()[-1]); r=d["roofline_step"]; e=d["e2e"]
    print("c5 value %.0f step %.3f ms k1 %.3f k3 %.3f frac %.3f launches %d" % (d["value"], d["ms_per_step"], r["k1_ms"], r["k3_ms"], d["roofline"]["frac"], d["gpu_launches"]))
    print("e2e %.0f (%.1f ms)  pipelined %.0f (%.1f ms, identical %s)  int16 pipelined %.0f (%.1f ms)" % (e["value"], e["ms_per_step"], e["pipelined"]["value"], e["pipelined"]["ms_per_step"], e["pipelined"]["outputs_identical"], e["int16_pcm_pipelined"]["value"], e["int16_pcm_pipelined"]["ms_per_step"]))
    print("link", e["per_rank_link_gbs"], e["pipelined"]["per_rank_link_gbs"], e["int16_pcm_pipelined"]["per_rank_link_gbs"])
    print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["cores"])
    for c in d.get("configs", []): print("  cfg", c)
except Exception as ex:
    print("bench parse failed", ex); print(open("gpurun_out/bench_default.err").read()[-1500:])
PY
timeout 600 ./benches/bench --c5 32 600 2>&1 | tail -3
timeout 300 ./benches/bench 2>&1 | tail -5
