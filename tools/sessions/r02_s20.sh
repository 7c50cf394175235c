#!/bin/bash
# round-2 session 20: short-K head GEMMs on one-tile CTAs (two per SM) instead of the persistent kernel
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for k in 0 3 5 0; do
  MICKEY_GEMM_PERSIST_MIN_KCHUNKS=$k python bench.py --steps 10 --warmup 3 --blocks 3 --no-cpu-baseline --no-eager-baseline --no-c2 2>&1 | tail -1 > gpurun_out/r02_s20_bench_k$k.json
  python - <<PY
import json
d = json.load(open("gpurun_out/r02_s20_bench_k$k.json"))
s = d["stage_ms"]
print("min_k=$k", round(d["value"], 1), round(d["ms_per_step"], 2), d["clocks"]["sm_mhz"], {k: s[k] for k in s if k.startswith("head.") or k in ("vit.qkv", "vit.attention")})
PY
done
MICKEY_GEMM_PERSIST_MIN_KCHUNKS=5 python -m pytest tests -q -m gpu -x -k "golden or parity or engine or gemm or conv" 2>&1 | tail -2
