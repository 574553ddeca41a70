#!/bin/bash
# Round-2 call T: apply and repair phases of the blocked kernel on 192 threads (4 columns x 2 atoms per item): suite, cycles, bench.
TAG=${1:-r02_t}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; head -4 $OUT/bcd_timing.log
timeout 600 python bench.py --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
timeout 120 python scripts/loop_trace.py device 4 2>&1 | tail -2
ls $OUT
