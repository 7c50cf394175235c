#!/bin/bash
# round-2 session 22: every 3rd pair of logits on the FMA-pipe exp2 (MICKEY_ATTN_POLY=3) -- parity and time
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MICKEY_ATTN_POLY=3 timeout -s KILL 300 python -m pytest tests -m gpu -x -q -k "golden or parity or attention" 2>&1 | tail -3
cp gpurun_out/parity_metrics.json gpurun_out/parity_metrics_poly3.json 2>/dev/null
MICKEY_ATTN_POLY=3 timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
for p in 4 3; do MICKEY_ATTN_POLY=$p timeout -s KILL 100 python tools/attn_bench.py 2>&1 | tail -1; done
for p in 3 4; do
  MICKEY_ATTN_POLY=$p python bench.py --steps 10 --warmup 3 --blocks 3 --no-cpu-baseline --no-eager-baseline --no-c2 2>&1 | tail -1 > gpurun_out/r02_s22_bench_p$p.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_s22_bench_p$p.json"))
print("poly=$p", round(d["value"], 1), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"], d["stage_ms"]["vit.attention"])
PY
done
