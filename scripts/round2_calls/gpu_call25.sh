#!/bin/bash
# round 2, call 25: resolve kernel flags 16-byte granules of confirmed end positions (wave chain 132 -> 84 steps), fewer atomics
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "search or Search" 2>&1 | tail -5 > gpurun_out/r02_c25_tests.txt
cat gpurun_out/r02_c25_tests.txt
rm -f gpurun_out/r02_c25_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c25_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 30 --warmup 3 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], 'e2e_ms', d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c25_bench.txt
}
run search_n32_h4096 TA_X=1
run search_n32_h4096 TA_QGRAM_QCAP=1
run search_n32_h4096 TA_SEARCH_FILTER=pigeon
run search_all_n32_h4096 TA_X=1
run search_affine_n32_h4096 TA_X=1
run search_n64_h4096 TA_X=1
cat gpurun_out/r02_c25_bench.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r02_launches_search_n32_h4096.csv \
    python bench.py --workload search_n32_h4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
grep -E "search_|memset" gpurun_out/r02_launches_search_n32_h4096.csv | tail -8 | cut -c1-60,180-330
ncu --set full --clock-control none --import-source on -k regex:'search_qgram' -s 6 -c 2 -f -o gpurun_out/prof_search_qgram_n32_h4096 \
    python bench.py --workload search_n32_h4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_search_qgram.log 2>&1
TA_TRACE_SEARCH=1 python bench.py --workload search_n32_h4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | grep "ta search" | tail -2
ncu --set full --clock-control none --import-source on -k regex:'search_wave' -s 3 -c 1 -f -o gpurun_out/prof_search_wave_n32_h4096 \
    python bench.py --workload search_n32_h4096 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/prof_search_wave.log 2>&1
