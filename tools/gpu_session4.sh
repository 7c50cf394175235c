#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run t_ops 600 python -m pytest tests/test_gpu_ops.py -q --timeout=120
run t_parity 900 python -m pytest tests/test_gpu_parity.py -q --timeout=400
run microbench 600 python tools/microbench.py
run bench 600 python bench.py --steps 20 --warmup 5
run ncu_gemm 900 ncu --set full --clock-control none -k regex:gemm_tc_kernel -c 4 -o gpurun_out/prof_gemm python tools/ncu_targets.py fc1 proj
run ncu_attn 900 ncu --set full --clock-control none -k regex:attention_tc -c 1 -o gpurun_out/prof_attn python tools/ncu_targets.py attention
run ncu_samp 900 ncu --set full --clock-control none -k regex:sampler_ -c 4 -o gpurun_out/prof_samp python tools/ncu_targets.py sampler
for f in t_ops t_parity; do echo "--- $f"; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/$f.log | cut -c1-400 | head -40; done
echo "--- microbench"; cat gpurun_out/microbench.log
echo "--- bench"; tail -n 2 gpurun_out/bench.log | cut -c1-4000
ls -la gpurun_out | head -30
