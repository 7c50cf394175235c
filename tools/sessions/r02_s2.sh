#!/bin/bash
# round-2 session 2: the new parity tests (full-size ViT-B / ViT-L fixtures, end-to-end pose, failure contract, flags)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-400)"; return $rc; }
run s2_parity 900 python -m pytest tests/test_gpu_parity.py -q -x
grep -E "^(FAILED|ERROR)|Error|assert" gpurun_out/s2_parity.log | head -40
run s2_parity_all 900 python -m pytest tests/test_gpu_parity.py -q
grep -E "^(FAILED|ERROR)" gpurun_out/s2_parity_all.log | head -40
cat gpurun_out/parity_metrics.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items(): print(k, {a:float('%.2e'%b) for a,b in v.items()})"
