#!/bin/bash
# One gpurun session: unit tests (SIMT first, then tcgen05), parity tests, a short bench, an ncu launch list.
# Full logs go to gpurun_out/; the tail printed here is what comes back on the console.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
run() {  # name, timeout, cmd...
  local name=$1 to=$2; shift 2
  timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1
  echo "== $name rc=$? :: $(tail -n 1 gpurun_out/$name.log | cut -c1-200)"
}
run t1_ops_simt 600 python -m pytest tests/test_gpu_ops.py -q -k "not tc] and not patch" --timeout=120
MICKEY_GEMM_IMPL=simt run t2_parity_simt 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
cp gpurun_out/parity_metrics.json gpurun_out/parity_metrics_simt.json 2>/dev/null
run t3_ops_tc 400 python -m pytest tests/test_gpu_ops.py -q -k "tc] or patch" --timeout=60
run t4_parity_tc 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
run smoke 300 python __graft_entry__.py --smoke
run bench 900 python bench.py --steps 10 --warmup 3
if [ "$1" == "ncu" ]; then
  run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline
fi
for f in t1_ops_simt t2_parity_simt t3_ops_tc t4_parity_tc; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$f.log | cut -c1-300 | head -40; done
echo "--- bench"; tail -n 3 gpurun_out/bench.log | cut -c1-3000
echo "--- smoke"; tail -n 3 gpurun_out/smoke.log | cut -c1-600
