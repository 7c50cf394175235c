#!/bin/bash
# round-1 profile artifacts: ncu launch list of the bench command (the timed latency steps only)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MICKEY_NCU_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_v3.csv \
  python bench.py --steps 3 --warmup 3 --depth 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
echo "launch list rc=$? rows=$(wc -l < gpurun_out/launches_v3.csv)"
