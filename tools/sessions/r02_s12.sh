#!/bin/bash
# round-2 session 12: sampler row-end masking without branches; attention straight-line path; C1-model workload
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 2 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s12_ops 600 python -m pytest tests/test_gpu_ops.py -q -k "sampler or attention"
run s12_micro 200 python tools/microbench.py
grep -E "sample" gpurun_out/s12_micro.log
timeout -s KILL 100 python tools/attn_bench.py
run s12_c1 600 python bench.py --workload c1 --steps 20 --warmup 5
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s12_c1.log").read().strip().splitlines()[-1])
print("c1 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "lat", round(d["latency_ms_single_step"], 3), d["clocks"]["sm_mhz"], "cpu", d.get("cpu_baseline"), "eager", d.get("gpu_eager_baseline"))
print(d["stage_ms"])
PY
run s12_bench 600 python bench.py --no-cpu-baseline --no-eager-baseline --no-c2 --blocks 3
python - <<'PY'
import json
d = json.loads(open("gpurun_out/s12_bench.log").read().strip().splitlines()[-1])
print("c3 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), d["clocks"]["sm_mhz"], "sampler", d["stage_ms"].get("solve.sample_outer"), "attn", d["stage_ms"].get("vit.attention"))
PY
