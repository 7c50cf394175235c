#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"; return $rc; }
run t_gpu 300 python -m pytest tests -q -m gpu --timeout=120
run microbench 120 python tools/microbench.py
run bench 200 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
run bench_c3 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1
MICKEY_GEMM_WIDE=0 run bench_c3_nowide 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_gpu.log | cut -c1-400 | head -40
echo "--- microbench"; head -8 gpurun_out/microbench.log
for f in bench bench_c3 bench_c3_nowide; do echo "--- $f"; tail -n 1 gpurun_out/$f.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','latency_ms_single_step')}, d['e2e']['value'], d.get('roofline_vit_gemm',{}).get('achieved')); print({k:v for k,v in list(d['stage_ms'].items())[:10]})"; done
