#!/bin/bash
# round-2 session 8: attention variants (FMA-pipe exp2 fraction x packed fp32x2), matcher one-tile TMA route, suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
for pk in 0 1; do for poly in 0 8 4; do
  MICKEY_ATTN_PACK2=$pk MICKEY_ATTN_POLY=$poly timeout -s KILL 100 python tools/attn_bench.py 2>&1 | sed "s/^/pack=$pk /"
done; done
run s8_tests 1200 python -m pytest tests -m gpu -q
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s8_tests.log | head -40
MICKEY_ATTN_PACK2=1 MICKEY_ATTN_POLY=0 run s8_attn_tests 300 python -m pytest tests/test_gpu_ops.py -q -k attention
run s8_micro 200 python tools/microbench.py
grep -E "match|matcher|sample" gpurun_out/s8_micro.log
for cfg in "0 4" "1 4" "1 0" "1 8"; do set -- $cfg
  MICKEY_ATTN_PACK2=$1 MICKEY_ATTN_POLY=$2 timeout -s KILL 300 python bench.py --no-cpu-baseline --no-eager-baseline --no-c2 --blocks 3 > gpurun_out/s8_bench_$1_$2.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/s8_bench_$1_$2.log").read().strip().splitlines()[-1])
print("pack=$1 poly=$2 value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "attn ms", d["stage_ms"].get("vit.attention"), "clocks", d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
PY
done
