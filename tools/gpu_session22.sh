#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout -s KILL 150 python -m pytest tests/test_gpu_ops.py -q --timeout=60 -k "residual_layernorm" -x 2>&1 | tail -8
echo "== ops rc=$?"
timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q --timeout=100 -x 2>&1 | tail -4
for v in 1 0; do MICKEY_FUSE_LN=$v timeout -s KILL 150 python bench.py --steps 30 --warmup 6 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(\"fuse=$v\", {k:round(d[k],4) for k in (\"value\",\"ms_per_step\",\"latency_ms_single_step\")}, round(d[\"e2e\"][\"value\"],1), d[\"gpu_launches\"], {k:d[\"stage_ms\"].get(k) for k in (\"vit.proj\",\"vit.fc2\",\"vit.layernorm\")})"; done
