#!/bin/bash
# round 2, call 28: q-gram scan only from 64 MB (8 MB for needles > 32) per call; full GPU suite, stress with the scan forced,
# sanitizers with the scan forced, bench lines
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_c28_tests.txt
cat gpurun_out/r02_c28_tests.txt
TA_SEARCH_FILTER=qgram timeout 300 python scripts/stress_search.py 90 11 2>&1 | tail -2 | tee gpurun_out/r02_c28_stress.txt
TA_SEARCH_FILTER=qgram TA_QGRAM_QCAP=3 timeout 300 python scripts/stress_search.py 45 12 2>&1 | tail -2 | tee -a gpurun_out/r02_c28_stress.txt
timeout 300 python scripts/stress_search.py 45 13 2>&1 | tail -2 | tee -a gpurun_out/r02_c28_stress.txt
timeout 900 compute-sanitizer --tool memcheck python scripts/sanitize_round2.py > gpurun_out/r02_c28_memcheck.log 2>&1; tail -2 gpurun_out/r02_c28_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck python scripts/sanitize_round2.py > gpurun_out/r02_c28_racecheck.log 2>&1; tail -2 gpurun_out/r02_c28_racecheck.log
rm -f gpurun_out/r02_c28_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c28_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 30 --warmup 3 --no-cpu-baseline --no-configs 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], 'e2e_ms', d['e2e']['ms_per_step'], d['gpu_launches'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c28_bench.txt
}
run search_n32_h4096 TA_X=1
run search_all_n32_h4096 TA_X=1
run search_affine_n32_h4096 TA_X=1
run search_n64_h4096 TA_X=1
run search_n32_h4096 TA_SEARCH_FILTER=pigeon
cat gpurun_out/r02_c28_bench.txt
python bench.py > gpurun_out/r02_c28_bench_default.json 2> gpurun_out/r02_c28_bench_default.err; tail -c 300 gpurun_out/r02_c28_bench_default.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c28_bench_ref.json 2> gpurun_out/r02_c28_bench_ref.err; tail -c 300 gpurun_out/r02_c28_bench_ref.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
