#!/bin/bash
# round-2 session 5: ncu --set full of the batch-shaped hot kernels (one report per target, summarised ON the box:
# only the text summaries and the small reports travel back), ncu launch list of the bench command (C3 latency steps)
mkdir -p gpurun_out profiles_tmp
export PYTHONUNBUFFERED=1 NCU_REPS=1
for t in dual_b fc1_b qkv_b fc2_b conv_b attention_b sampler; do
  timeout -s KILL 400 ncu --set full --clock-control none -f -o profiles_tmp/r02_$t python tools/ncu_targets.py $t > profiles_tmp/ncu_$t.log 2>&1
  echo "== ncu $t rc=$? $(ls -la profiles_tmp/r02_$t.ncu-rep 2>/dev/null | awk '{print $5}') bytes"
  python tools/ncu_summary.py profiles_tmp/r02_$t.ncu-rep > gpurun_out/r02_ncu_full_$t.txt 2>&1
  sz=$(stat -c %s profiles_tmp/r02_$t.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -lt 12000000 ]; then cp profiles_tmp/r02_$t.ncu-rep gpurun_out/; fi
done
MICKEY_NCU_RANGE=1 timeout -s KILL 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_launches_c3.csv python bench.py --steps 2 --warmup 3 --blocks 1 --no-cpu-baseline --no-eager-baseline --no-c2 > gpurun_out/s5_launches.log 2>&1
echo "launch rows: $(wc -l < gpurun_out/r02_launches_c3.csv)"
du -sh gpurun_out; ls -la gpurun_out | head -40
for t in dual_b fc1_b conv_b attention_b; do echo "######## $t"; grep -v "^$" gpurun_out/r02_ncu_full_$t.txt | head -150; done
