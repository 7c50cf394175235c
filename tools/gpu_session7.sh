#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
run() { local name=$1 to=$2; shift 2; timeout "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
run t_gpu 900 python -m pytest tests -q -m gpu --timeout=400
run microbench 600 python tools/microbench.py
run bench 600 python bench.py --steps 20 --warmup 6
run bench_ref 600 python bench.py --impl reference --steps 3 --warmup 1
run ncu_launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 6 --no-cpu-baseline
run ncu_gemm 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -c 2 -o gpurun_out/prof_gemm python tools/ncu_targets.py fc1
run ncu_attn 600 ncu --set full --clock-control none -k regex:attention_tc -c 1 -o gpurun_out/prof_attn python tools/ncu_targets.py attention
run ncu_dual 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 1 -c 1 -o gpurun_out/prof_dual python tools/ncu_targets.py dual
run ncu_conv 600 ncu --set full --clock-control none -k regex:gemm_tc_kernel -c 1 -o gpurun_out/prof_conv python tools/ncu_targets.py conv
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/t_gpu.log | cut -c1-400 | head -40
echo "--- microbench"; cat gpurun_out/microbench.log
echo "--- bench"; tail -n 1 gpurun_out/bench.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches','clocks','cpu_baseline')}); print(d['stage_ms'])"
echo "--- bench ref"; tail -n 1 gpurun_out/bench_ref.log | cut -c1-1500
ls -la gpurun_out
