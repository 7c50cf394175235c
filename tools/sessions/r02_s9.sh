#!/bin/bash
# round-2 session 9: attention tail skipping -- tests, attention timing, C3 bench; final-build ncu evidence (matcher with TMA
# stores, GEMMs with the 8-column epilogue, attention), launch list of the bench command
mkdir -p gpurun_out profiles_tmp
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s9_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s9_tests.log | head -30
timeout -s KILL 100 python tools/attn_bench.py
run s9_bench 600 python bench.py --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s9_bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "lat", round(d["latency_ms_single_step"], 3), "launches", d["gpu_launches"], d["clocks"])
for k in ("roofline", "roofline_step_tensor", "roofline_vit_gemm", "roofline_attention", "roofline_head_conv", "roofline_matcher", "roofline_matcher_pass2", "roofline_sampler", "roofline_solver"):
    r = d.get(k)
    if r: print("   ", k, r.get("kernel"), round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
print("    stage_ms", d["stage_ms"])
print("    latency_c2", {k: v for k, v in d["latency_c2"].items() if k != "workload"})
PY
export NCU_REPS=1
for t in dual_b fc1_b qkv_b fc2_b conv_b attention_b; do
  timeout -s KILL 400 ncu --set full --clock-control none -f -o profiles_tmp/r02_$t python tools/ncu_targets.py $t > profiles_tmp/ncu_$t.log 2>&1
  python tools/ncu_summary.py profiles_tmp/r02_$t.ncu-rep > gpurun_out/r02b_ncu_full_$t.txt 2>&1
done
MICKEY_NCU_RANGE=1 timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02b_launches_c3.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline --no-eager-baseline --no-c2 > gpurun_out/s9_launches.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/r02b_launches_c3.csv)"; du -sh gpurun_out
for t in dual_b fc1_b qkv_b conv_b attention_b; do echo "#### $t"; grep -A40 "mk::gemm_tc\|mk::attention_tc\|mk::matcher" gpurun_out/r02b_ncu_full_$t.txt | grep -E "Kernel Name|time_duration|dram__bytes|tensor_cycles_active.avg|issue_active|pipe_xu" ; done
