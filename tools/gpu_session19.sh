#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { tail -n 1 "$1" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],4) for k in ('value','ms_per_step','latency_ms_single_step')}, round(d['e2e']['value'],1)); print({k:v for k,v in list(d['stage_ms'].items())[:12]})"; }
timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/b_narrow1.log 2>&1; echo "== narrow=1 depth3"; show gpurun_out/b_narrow1.log
MICKEY_GEMM_NARROW=0 timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/b_narrow0.log 2>&1; echo "== narrow=0 depth3"; show gpurun_out/b_narrow0.log
timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --depth 4 > gpurun_out/b_d4.log 2>&1; echo "== depth4"; show gpurun_out/b_d4.log
timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline --depth 6 > gpurun_out/b_d6.log 2>&1; echo "== depth6"; show gpurun_out/b_d6.log
timeout 200 python -m pytest tests/test_gpu_ops.py -q --timeout=60 2>&1 | tail -2
