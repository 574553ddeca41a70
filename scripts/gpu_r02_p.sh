#!/bin/bash
# Round-2 call P: split-K of the [B_[:, subset] | C_] product (A/B on one box), whole GPU suite.
TAG=${1:-r02_p}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -3 $OUT/pytest_gpu.log
for V in "MODL_TC_SPLIT2=1" "MODL_TC_SPLIT2=0" "MODL_TC_SPLIT2=1"; do
  env $V timeout 300 python bench.py --no-cpu --no-e2e > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$V.json"))
print("$V", "value %.0f ms/step %.4f (min %.4f max %.4f)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"]), {k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
  env $V timeout 120 python scripts/loop_trace.py device 4 2>&1 | tail -2
done
ls $OUT
