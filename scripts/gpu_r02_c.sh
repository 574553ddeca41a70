#!/bin/bash
# Round-2 call C: pipelined norm exchange of the cluster dictionary kernel (parity + cycles per atom), A/B of the bounded
# mbarrier wait, timeline of the two-stream loop, bench.
TAG=${1:-r02_c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "update_dict" > $OUT/pytest_bcd.log 2>&1; echo "exit $?" >> $OUT/pytest_bcd.log; tail -5 $OUT/pytest_bcd.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
MODL_B200_LIB=$PWD/modl_b200/libmodl_b200_unbounded.so timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing_unbounded.log 2>&1; cat $OUT/bcd_timing_unbounded.log
timeout 120 python scripts/loop_trace.py device 10 > $OUT/trace_device.log 2>&1; cat $OUT/trace_device.log
timeout 120 python scripts/loop_trace.py pinned 10 > $OUT/trace_pinned.log 2>&1; cat $OUT/trace_pinned.log
timeout 900 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q > $OUT/pytest_df.log 2>&1; echo "exit $?" >> $OUT/pytest_df.log; tail -5 $OUT/pytest_df.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
MODL_B200_LIB=$PWD/modl_b200/libmodl_b200_unbounded.so timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_unbounded.json 2> $OUT/bench_unbounded.err
python - <<PY
import json
d=json.load(open("$OUT/bench_unbounded.json"))
print("unbounded wait: value %.0f ms/step %.4f" % (d["value"], d["ms_per_step"]), {k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
ls $OUT
