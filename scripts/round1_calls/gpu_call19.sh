#!/bin/bash
# end-of-round refresh: full GPU suite, every bench line (incl. the ragged workload), default + reference lines
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/c19_pytest.log 2>&1
tail -3 gpurun_out/c19_pytest.log
STEPS=100 bash scripts/bench_all.sh > gpurun_out/c19_bench_all.log 2>&1
tail -15 gpurun_out/c19_bench_all.log
python bench.py > gpurun_out/c19_bench_default.json 2> gpurun_out/c19_bench_default.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c19_bench_reference.json 2>&1
TA_BLK_DUO=0 STEPS=100 bash scripts/quick_bench.sh lev_k8_ragged96_160
