#!/bin/bash
# round-2 session 11: row-structured pitched sampler; compute-sanitizer memcheck of one small forward; suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s11_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s11_tests.log | head -30
run s11_micro 200 python tools/microbench.py
grep -E "sample|match" gpurun_out/s11_micro.log
run s11_memcheck 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -c "import __graft_entry__ as g; g.smoke()"
grep -E "ERROR SUMMARY|Invalid|smoke ok" gpurun_out/s11_memcheck.log | head
run s11_bench 600 python bench.py --no-cpu-baseline --no-eager-baseline --blocks 3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s11_bench.log").read().strip().splitlines()[-1])
print("value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), d["clocks"]["sm_mhz"], "sampler", d["stage_ms"].get("solve.sample_outer"), "attn", d["stage_ms"].get("vit.attention"), "c2", round(d["latency_c2"]["value"], 1))
PY
