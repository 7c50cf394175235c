#!/bin/bash
# round-2 session 10: single-wait TMEM load in attention, fixed-shift LSE partials; sensitivity of the ViT-L scores error
# to the attention variant; suite, microbench, bench
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s10_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s10_tests.log | head -30
for cfg in "0 4" "1 4" "1 8" "1 0"; do set -- $cfg
  MICKEY_ATTN_PACK2=$1 MICKEY_ATTN_POLY=$2 timeout -s KILL 200 python -m pytest tests/test_gpu_parity.py -q -k "golden and (vitl_720 or vitb_720 or vits_720)" > /dev/null 2>&1
  python - <<PY
import json
d = json.load(open("gpurun_out/parity_metrics.json"))
print("pack=$1 poly=$2", {k: {a: float("%.3g" % b) for a, b in v.items() if a in ("dsc", "scores", "final_scores", "depth")} for k, v in d.items() if k.endswith("720x540") and not k.startswith("pose")})
PY
done
timeout -s KILL 100 python tools/attn_bench.py
run s10_micro 200 python tools/microbench.py
grep -E "match|matcher" gpurun_out/s10_micro.log
run s10_bench 600 python bench.py --no-cpu-baseline --no-eager-baseline
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s10_bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "lat", round(d["latency_ms_single_step"], 3), d["clocks"])
for k in ("roofline_attention", "roofline_vit_gemm", "roofline_head_conv", "roofline_matcher", "roofline_matcher_pass2"):
    r = d.get(k)
    if r: print("   ", k, round(r["achieved"], 1), r["unit"], "frac", round(r["frac"], 3))
print("    stage_ms", d["stage_ms"])
print("    c2", d["latency_c2"]["value"], d["latency_c2"]["latency_ms_single_step"], d["latency_c2"]["stage_ms"])
PY
