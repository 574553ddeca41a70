#!/bin/bash
# Round-2 call M: TMA bulk rows in the blocked dictionary kernel, raw-X operand of the full-width B_ product (A/B).
TAG=${1:-r02_m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q > $OUT/pytest_kernels.log 2>&1; echo "exit $?" >> $OUT/pytest_kernels.log; tail -5 $OUT/pytest_kernels.log
timeout 900 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q > $OUT/pytest_df.log 2>&1; echo "exit $?" >> $OUT/pytest_df.log; tail -5 $OUT/pytest_df.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; head -4 $OUT/bcd_timing.log
for V in "MODL_TC_RAW_B=1" "MODL_TC_RAW_B=0"; do
  env $V timeout 600 python bench.py --no-cpu > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$V.json"))
print("$V", "value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
  tail -2 $OUT/bench_$V.err
done
timeout 120 python scripts/loop_trace.py device 6 > $OUT/trace_device.log 2>&1; tail -4 $OUT/trace_device.log
ls $OUT
