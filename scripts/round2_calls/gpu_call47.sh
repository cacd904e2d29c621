#!/bin/bash
# round 2, call 47: the full GPU suite and smoke() on the final commit
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/r02_c47_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
