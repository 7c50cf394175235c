#!/bin/bash
# round-2 session 18: cta_group::2 GEMM with 16 epilogue warps (four per TMEM lane quadrant) against 8
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MICKEY_GEMM_2SM_EPIWARPS=16 python -m pytest tests -q -m gpu -x -k "gemm or conv or linear" 2>&1 | tail -2
{
for w in 8 16 8 16; do MICKEY_GEMM_2SM_EPIWARPS=$w python tools/gemm_bench.py 2>&1 | sed "s/^/epiwarps=$w /"; done
} | tee gpurun_out/r02_s18_gemm.txt
MICKEY_GEMM_2SM_EPIWARPS=16 python -m pytest tests -q -m gpu -x -k "golden or parity or engine" 2>&1 | tail -2
MICKEY_GEMM_2SM_EPIWARPS=16 python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r02_s18_bench_c3_w16.json
python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r02_s18_bench_c3_w8.json
python - <<'PY'
import json
for t in ("w8", "w16"):
    d = json.load(open(f"gpurun_out/r02_s18_bench_c3_{t}.json"))
    print(t, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], d["roofline_vit_gemm"]["frac"], {k: d["stage_ms"][k] for k in ("vit.qkv", "vit.proj", "vit.fc1", "vit.fc2", "vit.attention", "head.att.qkv")}, d["latency_c2"]["value"])
PY
