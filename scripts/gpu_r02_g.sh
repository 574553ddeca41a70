#!/bin/bash
# Round-2 call G: per-phase cycles and ncu source-level capture of the block-wise dictionary kernel; full bench line;
# long-run drift numbers.
TAG=${1:-r02_g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:bcd_blocked -s 5 -c 1 \
    -f -o $OUT/prof_bcd_blocked python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_bcd_blocked.log 2>&1
ncu -i $OUT/prof_bcd_blocked.ncu-rep --page raw --csv > $OUT/prof_bcd_blocked.raw.csv 2>/dev/null
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
print("cpu", d["cpu_baseline"]["value"], "parity", d["parity"], "dense", d["extra"]["dense_codes"]["ms_per_step"])
PY
tail -3 $OUT/bench.err
timeout 120 python scripts/loop_trace.py pinned 8 > $OUT/trace_pinned.log 2>&1; tail -9 $OUT/trace_pinned.log
timeout 600 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q -s -k "long_run_drift" > $OUT/pytest_drift.log 2>&1; grep -i "drift\|passed\|failed" $OUT/pytest_drift.log | tail -4
ls -la $OUT
