#!/bin/bash
# round-2 session 14: FMA-pipe exp2 on packed fp32x2 pairs (MICKEY_ATTN_PACK2=2) against the scalar polynomial (=1);
# cta_group::2 GEMM ring depth 4 (64-column epilogue passes) against 5 (32-column passes)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for cfg in "1 4" "2 4" "2 3" "2 8" "2 2" "1 4" "2 4"; do
  set -- $cfg
  MICKEY_ATTN_PACK2=$1 MICKEY_ATTN_POLY=$2 python tools/attn_bench.py 2>&1 | sed "s/^/pack=$1 /"
done
} | tee gpurun_out/r02_s14_attn.txt
{
for st in 4 5 4 5; do MICKEY_GEMM_2SM_STAGES=$st python tools/gemm_bench.py 2>&1; done
} | tee gpurun_out/r02_s14_gemm.txt
MICKEY_GEMM_2SM_STAGES=5 python -m pytest tests -q -m gpu -x -k "gemm or conv" 2>&1 | tail -3
python -m pytest tests -q -m gpu -x -k "attention or golden or parity" 2>&1 | tail -5
python bench.py --steps 12 --warmup 4 2>&1 | tail -1 | tee gpurun_out/r02_s14_bench_c3.json
MICKEY_GEMM_2SM_STAGES=5 python bench.py --steps 12 --warmup 4 2>&1 | tail -1 | tee gpurun_out/r02_s14_bench_c3_st5.json
