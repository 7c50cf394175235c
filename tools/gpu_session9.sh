#!/bin/bash
mkdir -p gpurun_out
for d in 1 2 3 4; do
  timeout 300 python bench.py --steps 30 --warmup 6 --depth $d --no-cpu-baseline > gpurun_out/bench_d$d.log 2>&1
  echo "depth $d: $(tail -n 1 gpurun_out/bench_d$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','latency_ms_single_step')}, d['e2e']['value'])" 2>&1 | tail -n 1)"
done
