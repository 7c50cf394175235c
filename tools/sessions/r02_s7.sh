#!/bin/bash
# round-2 session 7: rewritten staged epilogue (8 columns per lane) -- op tests, full suite, microbench, C3 bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s7_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s7_tests.log | head -40
run s7_micro 200 python tools/microbench.py
cat gpurun_out/s7_micro.log
run s7_bench 600 python bench.py --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
for f in ("s7_bench",):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unparsable", e); continue
    print(f, "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "min/max", round(d["ms_per_step_min"], 3), round(d["ms_per_step_max"], 3),
          "e2e", round(d["e2e"]["value"], 1), "lat", round(d["latency_ms_single_step"], 3), "launches", d["gpu_launches"], "clocks", d["clocks"])
    for k in ("roofline", "roofline_step_tensor", "roofline_vit_gemm", "roofline_attention", "roofline_head_conv", "roofline_matcher", "roofline_matcher_pass2", "roofline_sampler", "roofline_solver"):
        r = d.get(k)
        if r: print("   ", k, r.get("kernel"), round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
    print("    stage_ms", d["stage_ms"])
    for k in ("latency_c2",):
        if k in d: print("   ", k, d[k])
PY
