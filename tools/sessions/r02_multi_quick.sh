#!/bin/bash
# closing 2-GPU check of the final build: NCCL ragged-shard check + a short C3 bench under torchrun
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 2 gpurun_out/$name.log | cut -c1-600)"; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run mq${N}_dist 200 $TR --master-port 29513 tools/dist_check.py || echo "DIST CHECK FAILED"
run mq${N}_bench 400 $TR --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 --blocks 3 --no-cpu-baseline --no-eager-baseline --no-c2
python - <<PY
import json
d = json.loads(open("gpurun_out/mq${N}_bench.log").read().strip().splitlines()[-1])
print("n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "per_rank", d.get("per_rank"), d["clocks"])
PY
