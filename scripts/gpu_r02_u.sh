#!/bin/bash
# Round-2 call U: fused apply_sub kernel validated (full GPU suite), end-to-end timeline (pinned rows) and copy diagnostics.
TAG=${1:-r02_u}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 120 python scripts/loop_trace.py pinned 16 > $OUT/trace_pinned.log 2>&1; tail -8 $OUT/trace_pinned.log | cut -c1-260
timeout 120 python scripts/loop_trace.py device 6 > $OUT/trace_device.log 2>&1; tail -3 $OUT/trace_device.log | cut -c1-260
timeout 200 python scripts/e2e_diag.py > $OUT/e2e_diag.log 2>&1; cat $OUT/e2e_diag.log
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv > $OUT/pcie.txt; cat $OUT/pcie.txt
