#!/bin/bash
mkdir -p gpurun_out
for d in 1 2; do
  timeout 600 python bench.py --workload c3 --steps 5 --warmup 4 --depth $d > gpurun_out/bench_c3_d$d.log 2>&1
  echo "c3 depth $d rc=$?: $(tail -n 1 gpurun_out/bench_c3_d$d.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','latency_ms_single_step','roofline_vit_gemm','roofline_attention','roofline_head_conv','roofline_matcher')}, d['e2e']['value'], d['stage_ms'])" 2>&1 | tail -n 1 | cut -c1-3000)"
done
tail -n 3 gpurun_out/bench_c3_d1.log | cut -c1-800
