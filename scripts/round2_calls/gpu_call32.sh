#!/bin/bash
# round 2, call 32: tiled duo kernel without extra shared memory (order in global scratch, counters borrowed from the clean tables)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_k_mutated or test_lev_duo_ragged_tiles or test_full_size_properties" 2>&1 | tail -3
rm -f gpurun_out/r02_c32_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c32_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 200 --warmup 5 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c32_bench.txt
}
run lev_k8_len128 TA_DUO_TILED=1
run lev_k8_len128 TA_DUO_TILED=0
run lev_k8_ragged96_160 TA_DUO_TILED=1
run lev_k8_ragged96_160 TA_DUO_TILED=0
run lev_k8_len128_R TA_DUO_TILED=1
run lev_k8_len128_R TA_DUO_TILED=0
cat gpurun_out/r02_c32_bench.txt
for t in 1 0; do
TA_DUO_TILED=$t ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:lev_bitpar -s 5 -c 2 --csv --log-file gpurun_out/r02_c32_ncu_ragged_tiled$t.csv \
    python bench.py --workload lev_k8_ragged96_160 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
grep lev_bitpar gpurun_out/r02_c32_ncu_ragged_tiled$t.csv | cut -d, -f5,13- | tail -6
done
