#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for d in 0 600 1000 1400; do MICKEY_ATTN_DEPHASE_CLKS=$d timeout 60 python tools/attn_bench.py 2>&1 | sed "s/^/dephase=$d /"; done
MICKEY_ATTN_DBG=2 timeout 60 python tools/attn_bench.py 2>&1 | tail -10
timeout 200 python -m pytest tests/test_gpu_ops.py -q --timeout=60 -k "sampler or attention" 2>&1 | tail -5
timeout 100 python tools/microbench.py 2>&1 | grep -i "sample\|attention"
