#!/bin/bash
# Round-2 call L: block-wise dictionary kernel v6 (sequential phases, Gram parts through shared memory, balanced 8-warp look-ahead behind the pipelined solver):
# all-gather exchange with st.async, norms from the solver): parity, cycles, bench.
TAG=${1:-r02_l}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "update_dict" > $OUT/pytest_bcd.log 2>&1; echo "exit $?" >> $OUT/pytest_bcd.log; tail -12 $OUT/pytest_bcd.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
tail -3 $OUT/bench.err
timeout 120 python scripts/loop_trace.py device 6 > $OUT/trace_device.log 2>&1; tail -6 $OUT/trace_device.log
timeout 120 python scripts/loop_trace.py pinned 8 > $OUT/trace_pinned.log 2>&1; tail -8 $OUT/trace_pinned.log
timeout 900 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q > $OUT/pytest_df.log 2>&1; echo "exit $?" >> $OUT/pytest_df.log; tail -5 $OUT/pytest_df.log
ls $OUT
