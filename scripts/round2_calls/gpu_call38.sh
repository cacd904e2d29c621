#!/bin/bash
# round 2, call 38: full GPU suite, smoke(), default bench line and reference arm on the final library
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r02_c38_tests.txt
cat gpurun_out/r02_c38_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r02_c38_bench_default.json 2> gpurun_out/r02_c38_bench_default.err; tail -c 200 gpurun_out/r02_c38_bench_default.json; echo
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_c38_bench_ref.json 2> gpurun_out/r02_c38_bench_ref.err; tail -c 200 gpurun_out/r02_c38_bench_ref.json; echo
TA_TRACE_SEARCH=1 python bench.py --workload search_n32_h4096 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-configs 2>&1 | grep "ta search" | tail -2
