#!/bin/bash
# round 2, call 23: full GPU suite on the new search path; scan variants (U=8 x 3 CTAs, U=4 x 5 CTAs); default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r02_c23_tests.txt
cat gpurun_out/r02_c23_tests.txt
rm -f gpurun_out/r02_c23_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c23_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 30 --warmup 3 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], 'e2e_ms', d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c23_bench.txt
}
run search_n32_h4096 TA_X=1
run search_n32_h4096 TA_QGRAM_U=8
run search_n32_h4096 TA_QGRAM_U=45
run search_n32_h4096 TA_QGRAM_U=8 TA_QGRAM_CTAS=2
run search_n64_h4096 TA_X=1
cat gpurun_out/r02_c23_bench.txt
python bench.py > gpurun_out/r02_c23_bench_default.json 2> gpurun_out/r02_c23_bench_default.err
tail -c 1500 gpurun_out/r02_c23_bench_default.json
