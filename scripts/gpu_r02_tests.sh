#!/bin/bash
# the whole GPU suite in one process (the order in which the tests reach the kernels matters: per-function attributes)
TAG=${1:-r02_tests}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -i "drift over\|passed\|failed\|^FAILED\|pytest exit" $OUT/pytest_gpu.log | tail -8 | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -x -q -p no:randomly > $OUT/pytest_gpu_x.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu_x.log; tail -3 $OUT/pytest_gpu_x.log
