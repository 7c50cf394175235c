#!/bin/bash
# round-2 session 6: tests of the TMA-store matcher + pitch-aware sampler, microbench, and a source-level ncu profile of the
# ViT-B qkv GEMM (why is the tensor pipe only 45 % active?)
mkdir -p gpurun_out profiles_tmp
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s6_ops 600 python -m pytest tests/test_gpu_ops.py -q -k "matcher or sampler"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s6_ops.log | head -30
run s6_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s6_tests.log | head -40
run s6_micro 200 python tools/microbench.py
grep -E "fill|copy|match|matcher|sample" gpurun_out/s6_micro.log
NCU_REPS=1 timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:gemm_tc -f -o profiles_tmp/qkv python tools/ncu_targets.py qkv_b > profiles_tmp/qkv.log 2>&1
ls -la profiles_tmp/qkv.ncu-rep
python tools/ncu_hot.py profiles_tmp/qkv.ncu-rep gemm_tc 60 > gpurun_out/r02_hot_qkv_b.txt 2>&1; head -80 gpurun_out/r02_hot_qkv_b.txt
NCU_REPS=1 timeout -s KILL 400 ncu --set full --import-source on --clock-control none -k regex:gemm_tc_persistent -f -o profiles_tmp/dual python tools/ncu_targets.py dual_b > profiles_tmp/dual.log 2>&1
python tools/ncu_summary.py profiles_tmp/dual.ncu-rep > gpurun_out/r02_ncu_full_dual_b_tma.txt 2>&1
python tools/ncu_hot.py profiles_tmp/dual.ncu-rep "gemm_tc_persistent_kernel<128, 7" 40 > gpurun_out/r02_hot_dual_b.txt 2>&1; head -60 gpurun_out/r02_hot_dual_b.txt
grep -E "Kernel Name|time_duration|dram__bytes|issue_active|stall" gpurun_out/r02_ncu_full_dual_b_tma.txt | head -60
