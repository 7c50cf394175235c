#!/bin/bash
# round-2 session 16: staged-epilogue variants of the cta_group::2 GEMM (MICKEY_GEMM_EPI_FLAGS: 1 = swizzled fp32 staging,
# 2 = fp16 staging for the fp16-store epilogue, 3 = both) -- is shared-memory traffic what the K = 768 GEMMs wait for?
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for f in 1 2; do MICKEY_GEMM_EPI_FLAGS=$f python -m pytest tests -q -m gpu -x -k "gemm or conv or linear" 2>&1 | tail -2; done
{
for f in 0 1 2 3 0 3; do MICKEY_GEMM_EPI_FLAGS=$f python tools/gemm_bench.py 2>&1 | sed "s/^/flags=$f /"; done
} | tee gpurun_out/r02_s16_gemm.txt
MICKEY_GEMM_EPI_FLAGS=3 python -m pytest tests -q -m gpu -x -k "golden or parity" 2>&1 | tail -2
