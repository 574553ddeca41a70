#!/bin/bash
# Round-2 call A: measurements the round-1 verdict asked for that need no new code:
#   knobs of the grid-wide dictionary update (flag barrier, columns per CTA) at the fMRI and image shapes,
#   per-atom cycle breakdown of the cluster kernel at cluster 16 vs 8, next-row throughputs with both recsys
#   bookkeeping modes, ncu --set full for bcd_update_kernel and the two recsys kernels.
TAG=${1:-r02_a}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; tail -3 $OUT/bcd_timing.log
for ARGS in "" "--flag-barrier" "--coop-min-cols 64" "--coop-min-cols 128" "--coop-min-cols 256" "--flag-barrier --coop-min-cols 128"; do
  N=$(echo "base $ARGS" | tr -d ' -')
  timeout 120 python scripts/phase_profile.py fmri image $ARGS --out $OUT/phases_$N.json > $OUT/phases_$N.log 2>&1
  echo "== $ARGS"; python - <<PY
import json
for o in json.load(open("$OUT/phases_$N.json")):
    print(o["shape"], "step %.3f ms" % o["wall_ms_per_step"], "bcd %.3f ms" % o["phase_ms_per_step"]["dict_bcd"])
PY
done
timeout 240 python scripts/next_rows_bench.py --budget 150 --out $OUT/next_rows.json > $OUT/next_rows.log 2>&1
tail -c 1200 $OUT/next_rows.log
timeout 200 ncu --set full --clock-control none --import-source on -k regex:bcd_update_kernel -s 4 -c 1 \
    -f -o $OUT/prof_bcd_update_fmri python scripts/phase_profile.py fmri > $OUT/ncu_bcd_update_fmri.log 2>&1
ncu -i $OUT/prof_bcd_update_fmri.ncu-rep --page raw --csv > $OUT/prof_bcd_update_fmri.raw.csv 2>/dev/null
for K in recsys_gram_dx_kernel recsys_update_B_kernel; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:$K -s 45 -c 1 \
      -f -o $OUT/prof_$K python scripts/next_rows_bench.py --only recsys --budget 60 > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
done
ls -la $OUT
