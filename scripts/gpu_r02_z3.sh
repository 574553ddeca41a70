#!/bin/bash
# quick A/B of a CD-kernel variant: its parity tests, then the bench line (device leg + dense-code leg)
TAG=${1:-r02_z3}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_dict_fact.py -m gpu -x -q -k "regression or cd_ or config2 or golden or oracle" > $OUT/pytest_cd.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_cd.log; tail -4 $OUT/pytest_cd.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
print("dense", d["extra"]["dense_codes"]["ms_per_step"])
PY
