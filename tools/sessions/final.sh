#!/bin/bash
# what the driver runs at round end, plus the microbenchmarks: smoke, GPU suite, bench (both arms)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 2 gpurun_out/$name.log | cut -c1-700)"; return $rc; }
run final_smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
run final_tests 1200 python -m pytest tests -m gpu -x -q
run final_micro 200 python tools/microbench.py
grep -E "sample|match" gpurun_out/final_micro.log
timeout -s KILL 100 python tools/attn_bench.py
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
