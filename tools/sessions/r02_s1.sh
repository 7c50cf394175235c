#!/bin/bash
# round-2 session 1: baseline suite, the two experimental kernels (first time on hardware), bench C2 / C3
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-400)"; return $rc; }
nvidia-smi -L
run s1_tests 600 python -m pytest tests -m gpu -x -q
MICKEY_TEST_EXPERIMENTAL=1 run s1_pp 180 python -m pytest tests/test_gpu_ops.py -q -k pingpong
MICKEY_TEST_EXPERIMENTAL=1 MICKEY_GEMM_2SM=1 run s1_2sm 180 python -m pytest tests/test_gpu_ops.py -q -k 2sm
ATTN_IMPL=1 run s1_attn1 120 python tools/attn_bench.py
ATTN_IMPL=3 run s1_attn3 120 python tools/attn_bench.py
run s1_micro 200 python tools/microbench.py
MICKEY_GEMM_2SM=1 run s1_micro2sm 200 python tools/microbench.py
run s1_c2 300 python bench.py --steps 30 --warmup 6 --no-cpu-baseline
run s1_c3 300 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 --no-cpu-baseline
for f in s1_pp s1_2sm s1_attn1 s1_attn3 s1_micro s1_micro2sm; do echo "---- $f"; tail -n 25 gpurun_out/$f.log | cut -c1-300; done
