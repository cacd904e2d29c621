#!/bin/bash
# refresh every bench line + launch lists for profiles/
mkdir -p gpurun_out
STEPS=100 bash scripts/bench_all.sh > gpurun_out/c7_bench_all.log 2>&1
tail -12 gpurun_out/c7_bench_all.log
python bench.py > gpurun_out/c7_bench_default.json 2> gpurun_out/c7_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c7_bench_reference.json 2>&1
for WL in lev_k8_len128 search_n32_h4096; do
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
done
python - <<'PY'
import json
d=json.load(open("gpurun_out/c7_bench_default.json")); print(d["ms_per_step"], d["pairs_per_s"], d["roofline"]["frac"], d["e2e"]["pairs_per_s"], d["cpu_baseline"])
PY
