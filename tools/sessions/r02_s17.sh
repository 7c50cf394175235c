#!/bin/bash
# round-2 session 17: cta_group::2 GEMM after dropping the GPU-scope fence from the accumulator hand-back
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
python -m pytest tests -q -m gpu -x -k "gemm or conv or linear" 2>&1 | tail -2
{
for f in 0 3 0; do MICKEY_GEMM_EPI_FLAGS=$f python tools/gemm_bench.py 2>&1 | sed "s/^/flags=$f /"; done
MICKEY_GEMM_2SM_STAGES=4 python tools/gemm_bench.py 2>&1 | sed "s/^/flags=0 /"
MICKEY_GEMM_2SM_STAGES=5 python tools/gemm_bench.py 2>&1 | sed "s/^/flags=0 /"
} | tee gpurun_out/r02_s17_gemm.txt
python -m pytest tests -q -m gpu -x -k "golden or parity or engine" 2>&1 | tail -2
python bench.py --steps 12 --warmup 4 2>&1 | tail -1 > gpurun_out/r02_s17_bench_c3.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_s17_bench_c3.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"], d["roofline_vit_gemm"]["frac"], d["stage_ms"])
print(d["latency_c2"]["value"], d["latency_c2"]["latency_ms_single_step"])
PY
