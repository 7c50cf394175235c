#!/bin/bash
# closing validation of round 2: what the driver runs (smoke, GPU suite, both bench arms), then evidence of the final build
# (memcheck of the kernels that changed last, launch list of the bench command, ncu --set full of the GEMMs whose epilogue changed)
mkdir -p gpurun_out profiles_tmp
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 2 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run final_smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
run final_tests 1200 python -m pytest tests -m gpu -x -q
run final_ref 400 python bench.py --impl reference --steps 3 --warmup 1
run final_bench 900 python bench.py
python - <<'PY'
import json
d = json.loads(open("gpurun_out/final_bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "min/max", round(d["ms_per_step_min"], 3), round(d["ms_per_step_max"], 3), "e2e", round(d["e2e"]["value"], 1), "lat", round(d["latency_ms_single_step"], 3), "launches", d["gpu_launches"], d["clocks"])
for k in ("roofline", "roofline_step_tensor", "roofline_vit_gemm", "roofline_attention", "roofline_head_conv", "roofline_matcher", "roofline_matcher_pass2", "roofline_sampler", "roofline_solver"):
    r = d.get(k)
    if r: print("   ", k, r.get("kernel"), round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3), "traffic", r.get("traffic"))
print("    stage_ms", d["stage_ms"])
for k in ("latency_c2", "gpu_eager_baseline", "cpu_baseline"):
    print("   ", k, d.get(k))
PY
run final_micro 200 python tools/microbench.py
run final_gemm 100 python tools/gemm_bench.py
timeout -s KILL 100 python tools/attn_bench.py | tee gpurun_out/final_attn.log
run final_memcheck 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_ops.py -q -x -k "gemm_2sm or matcher or attention or sampler"
MICKEY_NCU_RANGE=1 timeout -s KILL 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02c_launches_c3.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline --no-eager-baseline --no-c2 > gpurun_out/final_launches.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/r02c_launches_c3.csv)"
export NCU_REPS=1
for t in fc1_b qkv_b; do
  timeout -s KILL 300 ncu --set full --clock-control none -f -o profiles_tmp/r02c_$t python tools/ncu_targets.py $t > profiles_tmp/ncu_$t.log 2>&1
  python tools/ncu_summary.py profiles_tmp/r02c_$t.ncu-rep > gpurun_out/r02c_ncu_full_$t.txt 2>&1
  grep -E "Kernel Name|time_duration|tensor_cycles_active.avg|issue_active|dram__bytes" gpurun_out/r02c_ncu_full_$t.txt | head -8
done
du -sh gpurun_out
