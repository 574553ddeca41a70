#!/bin/bash
# Round-2 call F: block-wise dictionary kernel v2 (parity, per-phase cycles, bench A/B) and the ncu captures the round-1
# verdict listed as missing (bcd_update_kernel at the fMRI shape, the two recsys kernels), next-row throughputs.
TAG=${1:-r02_f}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "update_dict" > $OUT/pytest_bcd.log 2>&1; echo "exit $?" >> $OUT/pytest_bcd.log; tail -6 $OUT/pytest_bcd.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
timeout 120 python scripts/loop_trace.py device 6 > $OUT/trace_device.log 2>&1; tail -6 $OUT/trace_device.log
timeout 600 python bench.py --no-cpu --no-e2e > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
tail -3 $OUT/bench.err
if [ -z "$SKIP_NEXT" ]; then
timeout 300 python scripts/next_rows_bench.py --budget 200 --out $OUT/next_rows.json > $OUT/next_rows.log 2>&1
tail -c 1500 $OUT/next_rows.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bcd_update_kernel -s 4 -c 1 \
    -f -o $OUT/prof_bcd_update_fmri python scripts/phase_profile.py fmri > $OUT/ncu_bcd_update_fmri.log 2>&1
ncu -i $OUT/prof_bcd_update_fmri.ncu-rep --page raw --csv > $OUT/prof_bcd_update_fmri.raw.csv 2>/dev/null
for K in recsys_gram_dx_kernel recsys_update_B_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 45 -c 1 \
      -f -o $OUT/prof_$K python scripts/next_rows_bench.py --only recsys --budget 60 > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
done
fi
ls -la $OUT
