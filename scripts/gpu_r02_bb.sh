#!/bin/bash
# blocked dictionary kernel variant: parity tests, per-phase cycles, bench (device leg only)
TAG=${1:-r02_bb}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dict_fact.py -m gpu -x -q -k "update_dict or config2 or golden or reproducible" > $OUT/pytest_bcd.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_bcd.log; tail -3 $OUT/pytest_bcd.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; head -6 $OUT/bcd_timing.log | cut -c1-300
timeout 600 python bench.py --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
