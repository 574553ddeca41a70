#!/bin/bash
# Round-2 call E: block-wise coefficient-space dictionary kernel (parity, cycles per block), call fence off for final
# device rows (timeline), bench A/B against the per-atom kernels.
TAG=${1:-r02_e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "update_dict" > $OUT/pytest_bcd.log 2>&1; echo "exit $?" >> $OUT/pytest_bcd.log; tail -12 $OUT/pytest_bcd.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
timeout 120 python scripts/loop_trace.py device 8 > $OUT/trace_device.log 2>&1; tail -9 $OUT/trace_device.log
timeout 120 python scripts/loop_trace.py pinned 8 > $OUT/trace_pinned.log 2>&1; tail -9 $OUT/trace_pinned.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
print("dense", d["extra"]["dense_codes"]["ms_per_step"])
PY
tail -3 $OUT/bench.err
for V in "MODL_BCD_BLOCKED=0" "MODL_BCD_BLOCKED=0 MODL_BCD_PIPELINE=0"; do
  N=$(echo "$V" | tr -d ' =')
  env $V timeout 300 python bench.py --no-cpu --no-e2e > $OUT/bench_$N.json 2> $OUT/bench_$N.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$N.json"))
print("$V", "value %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], {k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
done
timeout 900 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q > $OUT/pytest_df.log 2>&1; echo "exit $?" >> $OUT/pytest_df.log; tail -5 $OUT/pytest_df.log
ls $OUT
