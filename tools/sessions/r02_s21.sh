#!/bin/bash
# round-2 session 21: the solver as ONE launch (gather + hypotheses + argmax + refinement; last-block-per-pair finalize)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 3 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
run s21_solver 300 python -m pytest tests/test_gpu_parity.py -q -x -k "solver or pose or failure or planted"
grep -E "^(FAILED|ERROR)|^E  " gpurun_out/s21_solver.log | head -20
run s21_tests 600 python -m pytest tests -m gpu -x -q
run s21_memcheck 300 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -k "solver_with_injected or failure"
run s21_racecheck 300 compute-sanitizer --tool racecheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -q -x -k "failure"
run s21_smoke 200 python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 12 --warmup 4 --no-cpu-baseline --no-eager-baseline 2>&1 | tail -1 > gpurun_out/r02_s21_bench_c3.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_s21_bench_c3.json"))
print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["clocks"]["sm_mhz"], d["gpu_launches"], {k: d["stage_ms"][k] for k in ("solve.ransac", "solve.sample_outer")})
c = d["latency_c2"]
print(c["value"], c["latency_ms_single_step"], c["gpu_launches"], {k: c["stage_ms"][k] for k in ("solve.ransac", "solve.sample_outer")})
PY
