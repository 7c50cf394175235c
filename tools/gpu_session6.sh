#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run t_ops 600 python -m pytest tests/test_gpu_ops.py -q --timeout=120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
run microbench 600 python tools/microbench.py
run bench 600 python bench.py --steps 20 --warmup 6
MICKEY_PDL=0 run bench_nopdl 600 python bench.py --steps 20 --warmup 6 --no-cpu-baseline
for f in t_ops t_parity; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/$f.log | cut -c1-400 | head -40; done
echo "--- microbench"; cat gpurun_out/microbench.log
echo "--- bench"; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')}); print(d['stage_ms'])"
echo "--- bench no PDL"; tail -n 1 gpurun_out/bench_nopdl.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e')})"
