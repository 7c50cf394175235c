#!/bin/bash
mkdir -p gpurun_out
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; echo "== $name rc=$? :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; }
run t_ops 600 python -m pytest tests/test_gpu_ops.py -q --timeout=120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
run debug_vitb 300 python tools/debug_stages.py vitb
run smoke 300 python __graft_entry__.py --smoke
run bench 600 python bench.py --steps 20 --warmup 5
run ncu_full 1200 ncu --set full --clock-control none --import-source on -c 40 -o gpurun_out/prof_r1b python tools/ncu_targets.py fc1 proj attention conv dual sampler linattn
for f in t_ops t_parity; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/$f.log | cut -c1-300 | head -40; done
echo "--- debug"; tail -n 25 gpurun_out/debug_vitb.log | cut -c1-400
echo "--- bench"; tail -n 2 gpurun_out/bench.log | cut -c1-6000
