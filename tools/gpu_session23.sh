#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in 200 150 100; do MICKEY_GEMM_WIDE_MIN_PCT=$v timeout -s KILL 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"wide_min=$v\", {k:round(d[k],4) for k in (\"value\",\"ms_per_step\",\"latency_ms_single_step\")}, round(d[\"e2e\"][\"value\"],1), {k:d[\"stage_ms\"].get(k) for k in (\"head.conv3x3\",\"head.conv1x1\",\"solve.ransac\")})"; done
for v in 200 150; do MICKEY_GEMM_WIDE_MIN_PCT=$v timeout -s KILL 200 python bench.py --workload c3 --steps 5 --warmup 4 --depth 1 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"c3 wide_min=$v\", {k:round(d[k],4) for k in (\"value\",\"ms_per_step\",\"latency_ms_single_step\")}, {k:d[\"stage_ms\"].get(k) for k in (\"head.conv3x3\",\"head.conv1x1\",\"vit.fc1\",\"vit.qkv\")})"; done
