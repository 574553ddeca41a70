#!/bin/bash
# Round-2 call J: shared-memory / shuffle micro-benchmarks, blocked kernel with release-time stamps.
TAG=${1:-r02_j}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 120 ./scripts/ubench/smem_ubench > $OUT/smem_ubench.log 2>&1; cat $OUT/smem_ubench.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; head -5 $OUT/bcd_timing.log
ls $OUT
