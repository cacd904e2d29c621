#!/bin/bash
# final confirmation on the last commit of the round: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c21_pytest.log 2>&1
tail -3 gpurun_out/c21_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 100 > gpurun_out/c21_bench_default.json 2> gpurun_out/c21_bench_default.err; tail -c 400 gpurun_out/c21_bench_default.json
