#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 60 python tools/attn_bench.py 2>&1 | tail -4
timeout 300 python -m pytest tests/test_gpu_ops.py -q --timeout=60 2>&1 | tail -5
timeout 100 python tools/microbench.py 2>&1 | grep -i "sample\|attention"
timeout 200 python -m pytest tests/test_gpu_parity.py -q --timeout=100 2>&1 | tail -5
timeout 200 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','latency_ms_single_step')}, d['e2e']['value']); print({k:v for k,v in list(d['stage_ms'].items())[:10]})"
timeout 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 > gpurun_out/bench_c3.log 2>&1; tail -n 1 gpurun_out/bench_c3.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','latency_ms_single_step')}, d['e2e']['value']); print({k:v for k,v in list(d['stage_ms'].items())[:10]})"
