#!/bin/bash
# 8-GPU check: C4 (B = 256 sharded 32 pairs per GPU) bench line and the config-5 submission run on 8 ranks
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 1 gpurun_out/$name.log | cut -c1-300)"; return $rc; }
nvidia-smi -L | wc -l
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
run m8_bench 500 $TR --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 4 --blocks 3
python tools/make_synthetic_mapfree.py --root /tmp/mf --split val --scenes 16 --queries 60 > /dev/null
run m8_subm_u8 400 $TR --master-port 29514 tools/run_submission.py --variant vitb --data_root /tmp/mf --split val --uint8 -o gpurun_out/subm8
python - <<PY
import json
d = json.loads(open("gpurun_out/m8_bench.log").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "min/max", round(d["ms_per_step_min"], 3), round(d["ms_per_step_max"], 3), "e2e", round(d["e2e"]["value"], 1), "e2e ms", round(d["e2e"]["ms_per_step"], 3), "h2d", d["e2e"]["h2d_bytes_per_step"])
print("per_rank", d.get("per_rank")); print(d["clocks"]); print(d["config"]["workload"])
PY
rm -rf gpurun_out/subm8
