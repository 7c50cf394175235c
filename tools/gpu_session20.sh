#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
show() { tail -n 1 "$1" | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:round(d[k],4) for k in ('value','ms_per_step','latency_ms_single_step')}, round(d['e2e']['value'],1)); print({k:v for k,v in list(d['stage_ms'].items())[:12]})"; }
timeout 300 python -m pytest tests -q -m gpu --timeout=120 > gpurun_out/t_gpu.log 2>&1; tail -3 gpurun_out/t_gpu.log
timeout 100 python __graft_entry__.py --smoke 2>&1 | tail -2
timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/bench.log 2>&1; echo "== c2"; show gpurun_out/bench.log
MICKEY_PDL=0 timeout 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline > gpurun_out/bench_nopdl.log 2>&1; echo "== c2 nopdl"; show gpurun_out/bench_nopdl.log
timeout 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 > gpurun_out/bench_c3.log 2>&1; echo "== c3"; show gpurun_out/bench_c3.log
MICKEY_PDL=0 timeout 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 > gpurun_out/bench_c3_nopdl.log 2>&1; echo "== c3 nopdl"; show gpurun_out/bench_c3_nopdl.log
