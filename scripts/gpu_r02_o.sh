#!/bin/bash
# Round-2 call O: same-box A/B of the blocked dictionary kernel variants (solver: pipelined | plain; Gram parts: shared memory | shuffles),
# raw-B GEMM with prefetching converters, BASELINE configs[4] at its stated size.
TAG=${1:-r02_o}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
for V in default plain shfl both default; do
  if [ "$V" = default ]; then unset MODL_B200_LIB; else export MODL_B200_LIB=$PWD/modl_b200/libmodl_b200_$V.so; fi
  timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing_$V.log 2>&1; sed -n 2,4p $OUT/bcd_timing_$V.log
  timeout 300 python bench.py --no-cpu --no-e2e > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$V.json"))
print("$V", "value %.0f ms/step %.4f (min %.4f max %.4f)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"]), {k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
done
unset MODL_B200_LIB
timeout 120 python scripts/loop_trace.py device 6 > $OUT/trace_device.log 2>&1; tail -3 $OUT/trace_device.log
timeout 900 python scripts/recsys_scale.py --rows 1000000 --out $OUT/recsys_scale.json > $OUT/recsys_scale.log 2>&1; tail -c 1500 $OUT/recsys_scale.log
ls $OUT
