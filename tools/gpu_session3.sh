#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run t_attn 200 python -m pytest tests/test_gpu_ops.py -q -k attention --timeout=60
if [ $? -ne 0 ]; then echo "!! tcgen05 attention failed: falling back to MICKEY_ATTN_IMPL=mma for the remaining steps"; export MICKEY_ATTN_IMPL=mma; fi
run t_ops 600 python -m pytest tests/test_gpu_ops.py -q --timeout=120 -k "not attention"
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
run smoke 300 python __graft_entry__.py --smoke
run bench 600 python bench.py --steps 20 --warmup 5
run ncu_gemm 900 ncu --set full --clock-control none -c 4 -o gpurun_out/prof_gemm python tools/ncu_targets.py fc1 proj
run ncu_attn 900 ncu --set full --clock-control none -c 2 -o gpurun_out/prof_attn python tools/ncu_targets.py attention
for f in t_attn t_ops t_parity; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/$f.log | cut -c1-400 | head -60; done
echo "--- bench"; tail -n 2 gpurun_out/bench.log | cut -c1-6000
ls -la gpurun_out | head -30
