#!/bin/bash
# round-2 session 13: source-level hot spots of the attention kernel at batch; sampler register cap check
mkdir -p gpurun_out profiles_tmp
export PYTHONUNBUFFERED=1 NCU_REPS=1
python tools/microbench.py 2>&1 | grep -E "sample"
timeout -s KILL 500 ncu --set full --import-source on --clock-control none -k regex:attention_tc -f -o profiles_tmp/attn python tools/ncu_targets.py attention_b > profiles_tmp/attn.log 2>&1
ls -la profiles_tmp/attn.ncu-rep
python tools/ncu_hot.py profiles_tmp/attn.ncu-rep attention_tc 120 > gpurun_out/r02_hot_attention_b.txt 2>&1
head -150 gpurun_out/r02_hot_attention_b.txt | cut -c1-170
