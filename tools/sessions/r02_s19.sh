#!/bin/bash
# round-2 session 19: packed fp32x2 GELU in the staged epilogue + 16 epilogue warps by default for GELU / residual epilogues
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests -q -m gpu -x -k "gemm or conv or linear" 2>&1 | tail -2
python tools/gemm_bench.py 2>&1 | tee gpurun_out/r02_s19_gemm.txt
python tools/microbench.py 2>&1 | grep -E "vit\.|gemm" | tee -a gpurun_out/r02_s19_gemm.txt
python -m pytest tests -q -m gpu -x -k "golden or parity or engine" 2>&1 | tail -2
python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r02_s19_bench_c3.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_s19_bench_c3.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["roofline_vit_gemm"]["frac"], d["stage_ms"])
print(d["latency_c2"]["value"], d["latency_c2"]["latency_ms_single_step"], d["latency_c2"]["stage_ms"])
PY
