#!/bin/bash
# multi-GPU session (N = number of GPUs of the box): NCCL check of the ragged sharded forward, C3/C4 bench, config-5 submission run
N=${1:-2}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1 to=$2; shift 2; timeout -s KILL "$to" "$@" > "gpurun_out/$name.log" 2>&1; local rc=$?; echo "== $name rc=$rc :: $(tail -n 2 gpurun_out/$name.log | cut -c1-900)"; return $rc; }
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
run m${N}_dist 300 $TR --master-port 29513 tools/dist_check.py || echo "DIST CHECK FAILED"
run m${N}_bench 600 $TR --master-port 29511 bench.py --gpus $N --steps 10 --warmup 4 --blocks 3
run m${N}_bench1 400 python bench.py --steps 10 --warmup 4 --blocks 3 --no-cpu-baseline --no-eager-baseline --no-c2
python tools/make_synthetic_mapfree.py --root /tmp/mf --split val --scenes 8 --queries 36 > /dev/null
run m${N}_subm_u8 600 $TR --master-port 29514 tools/run_submission.py --variant vitb --data_root /tmp/mf --split val --uint8 -o gpurun_out/subm_u8
run m${N}_subm_f32 600 $TR --master-port 29515 tools/run_submission.py --variant vitb --data_root /tmp/mf --split val -o gpurun_out/subm_f32
python - <<PY
import json, zipfile
for f in ("m${N}_bench", "m${N}_bench1"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.log").read().strip().splitlines()[-1])
        print(f, "n_gpus", d["n_gpus"], "value", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "h2d", d["e2e"]["h2d_bytes_per_step"], "per_rank", d.get("per_rank"), d["clocks"])
    except Exception as e:
        print(f, "unparsable", e)
for f in ("subm_u8", "subm_f32"):
    try:
        z = zipfile.ZipFile(f"gpurun_out/{f}/submission.zip"); n = z.namelist()
        print(f, len(n), "scene files; first line:", z.read(n[0]).decode().splitlines()[0])
    except Exception as e:
        print(f, "no zip", e)
PY
rm -rf gpurun_out/subm_u8 gpurun_out/subm_f32
