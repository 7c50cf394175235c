#!/bin/bash
# round-1 final validation: what the driver runs at round end
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout -s KILL 300 python bench.py > gpurun_out/final_bench.log 2>&1; echo "bench rc=$?"; tail -n 1 gpurun_out/final_bench.log | cut -c1-400
timeout -s KILL 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_ref.log 2>&1; echo "ref rc=$?"; tail -n 1 gpurun_out/final_bench_ref.log | cut -c1-400
timeout -s KILL 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 --no-cpu-baseline > gpurun_out/final_bench_c3.log 2>&1; echo "c3 rc=$?"; tail -n 1 gpurun_out/final_bench_c3.log | cut -c1-300
