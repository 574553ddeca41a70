#!/bin/bash
# Round-2 closing call at HEAD (no ncu: profiles/r02_final4_* hold the captures): smoke, the whole GPU suite, bench line.
TAG=${1:-r02_final5}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -i "drift over\|passed\|failed\|error" $OUT/pytest_gpu.log | tail -4 | cut -c1-300
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
print("cpu", d["cpu_baseline"]["value"], "parity", d["parity"], "dense", d["extra"]["dense_codes"]["ms_per_step"], "launches", d["gpu_launches"], "roofline frac", d["roofline"]["frac"])
PY
tail -3 $OUT/bench.err
