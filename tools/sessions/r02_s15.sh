#!/bin/bash
# round-2 session 15: cta_group::2 GEMM with the direct (thread == row, 256-bit stores, no staging) epilogue and a 6 / 7
# deep ring, against the staged epilogue with 4 stages
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for st in 6 7; do MICKEY_GEMM_2SM_STAGES=$st python -m pytest tests -q -m gpu -x -k "gemm or conv or linear" 2>&1 | tail -3; done
{
for st in 4 6 7 4 6; do MICKEY_GEMM_2SM_STAGES=$st python tools/gemm_bench.py 2>&1; done
} | tee gpurun_out/r02_s15_gemm.txt
MICKEY_GEMM_2SM_STAGES=6 python -m pytest tests -q -m gpu -x -k "golden or parity or engine" 2>&1 | tail -3
MICKEY_GEMM_2SM_STAGES=6 python bench.py --steps 12 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_s15_bench_c3_st6.json
python bench.py --steps 12 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_s15_bench_c3_st4.json
MICKEY_GEMM_2SM_STAGES=7 python bench.py --steps 12 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_s15_bench_c3_st7.json
python - <<'PY'
import json
for t in ("st4", "st6", "st7"):
    try:
        d = json.load(open(f"gpurun_out/r02_s15_bench_c3_{t}.json"))
        print(t, d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"], {k: d["stage_ms"][k] for k in ("vit.qkv", "vit.proj", "vit.fc1", "vit.fc2", "vit.attention", "head.att.qkv")})
    except Exception as e:
        print(t, "failed", e)
PY
