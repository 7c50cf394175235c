#!/bin/bash
# round-2 session 4: full GPU suite after the fixes, write-bandwidth probe, ncu --set full of the batch-shaped hot kernels,
# ncu launch list of the bench command (latency steps of the C3 workload)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s4_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)" gpurun_out/s4_tests.log | head -40
grep -E "^E  " gpurun_out/s4_tests.log | head -40
run s4_micro 200 python tools/microbench.py
grep -E "fill|copy|match|matcher" gpurun_out/s4_micro.log
run s4_ncu_full 900 ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_full python tools/ncu_targets.py dual_b fc1_b qkv_b fc2_b conv_b attention_b sampler
MICKEY_NCU_RANGE=1 run s4_launches 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline --no-eager-baseline --no-c2
echo "launch rows: $(wc -l < gpurun_out/r02_launches_c3.csv)"; ls -la gpurun_out/*.ncu-rep
cat gpurun_out/parity_metrics.json | python -c "
import json,sys
d=json.load(sys.stdin)
for k,v in d.items():
    if k.startswith('pose_e2e') or k.startswith('flags'): print(k, {a:float('%.2e'%b) for a,b in v.items()})"
