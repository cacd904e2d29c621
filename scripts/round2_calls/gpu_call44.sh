#!/bin/bash
# round 2, call 44: tile-ordered one-pair-per-thread kernel (ragged batches, k = 16 / transpositions): parity and timing
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_lev_duo_ragged_tiles or test_lev_k_mutated or tile-ordered or no-tile-ordering or test_length_hint" 2>&1 | tail -3
rm -f gpurun_out/r02_c44_bench.txt
run() { # name env...
  echo "== $*" >> gpurun_out/r02_c44_bench.txt
  env "${@:2}" timeout 300 python bench.py --workload $1 --steps 100 --warmup 5 --no-cpu-baseline --no-configs --no-e2e 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['pairs_per_s'], d['roofline']['frac'], d['parity_ok'], d['roofline']['kernel'])
except Exception as ex: print('ERR', ex)" >> gpurun_out/r02_c44_bench.txt
}
run lev_k16_ragged96_160 TA_X=1
run lev_k16_ragged96_160 TA_BLK_TILED=0
run lev_k16_len128 TA_BLK_TILED=1
run lev_k16_len128 TA_X=1
run rdamerau_k16_len512 TA_BLK_TILED=1
run rdamerau_k16_len512 TA_X=1
cat gpurun_out/r02_c44_bench.txt
