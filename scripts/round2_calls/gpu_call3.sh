#!/bin/bash
# round 2, call 3: ncu captures of the first diagonal-extension kernel (why is it not faster than the band kernels?)
mkdir -p gpurun_out
for wl in lev_k16_len4096 exp_len1024; do
TA_FR=1 ncu --set full --clock-control none --import-source on -k regex:'lev_fr' -s 3 -c 1 -f -o gpurun_out/prof_fr1_${wl} \
    python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/prof_fr1_${wl}.log 2>&1
done
for ctas in 6 16; do
  echo "== lev_k16_len4096 TA_FR=1 TA_FR_CTAS=$ctas" >> gpurun_out/r02_c3_bench.txt
  TA_FR=1 TA_FR_CTAS=$ctas timeout 300 python bench.py --workload lev_k16_len4096 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])" >> gpurun_out/r02_c3_bench.txt
done
cat gpurun_out/r02_c3_bench.txt
ls -la gpurun_out/*.ncu-rep
