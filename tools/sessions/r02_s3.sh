#!/bin/bash
# round-2 session 3: whole GPU suite (matcher v2, 2SM default, io, drop-in, parity), microbench, new bench (c3 default)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s3_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)" gpurun_out/s3_tests.log | head -40
grep -E "^E  " gpurun_out/s3_tests.log | head -60
run s3_micro 200 python tools/microbench.py
tail -n 22 gpurun_out/s3_micro.log
run s3_bench 600 python bench.py
MICKEY_GEMM_2SM=0 run s3_bench_no2sm 400 python bench.py --no-cpu-baseline --no-eager-baseline --no-c2
python - <<'PY'
import json
for f in ("s3_bench", "s3_bench_no2sm"):
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
    for k in ("latency_c2", "gpu_eager_baseline", "cpu_baseline"):
        if k in d: print("   ", k, d[k])
PY
