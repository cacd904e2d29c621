#!/bin/bash
# per-launch device times of one bench workload (ncu, cold-cache/serialised: compare shares)
WL=${1:-lev_k8_len128}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${WL}.csv \
    python bench.py --workload $WL --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/launches_${WL}.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/launches_${WL}.csv")) if len(r)>5 and r[0].isdigit()]
for r in rows[-12:]:
    print(r[4][:70], r[7], r[8], r[-1], "ns")
PY
