#!/bin/bash
# Round-2 evidence in ONE gpurun call (about 6-8 minutes of box time):
#   gpurun --timeout 900 -- 'bash scripts/gpu_round2.sh r02_v1'
# = scripts/gpu_round.sh (smoke, GPU tests, bench + reference arm, launch list, ncu --set full of the three
# dominant kernels of the config-2 step) plus the "next" rows: throughput at the configs[2..4] shapes, per-phase
# times, and ncu --set full captures of the kernels those rows add (grid-wide dictionary update with the L1
# projection at the fMRI shape, the recsys Gram / B_ recurrence kernels).
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
bash scripts/gpu_round.sh $TAG
timeout 200 python scripts/next_rows_bench.py --budget 120 --out $OUT/next_rows.json > $OUT/next_rows.log 2>&1
tail -c 1500 $OUT/next_rows.log
timeout 100 python scripts/phase_profile.py --out $OUT/next_rows_phases.json > $OUT/phase_profile.log 2>&1
tail -3 $OUT/phase_profile.log
if [ -z "$SKIP_NCU" ]; then
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:bcd_update_kernel -s 4 -c 1 \
      -f -o $OUT/prof_bcd_update_fmri python scripts/phase_profile.py fmri > $OUT/ncu_bcd_update_fmri.log 2>&1
  ncu -i $OUT/prof_bcd_update_fmri.ncu-rep --page raw --csv > $OUT/prof_bcd_update_fmri.raw.csv 2>/dev/null
  for K in recsys_gram_dx_kernel recsys_update_B_kernel; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 45 -c 1 \
        -f -o $OUT/prof_$K python scripts/next_rows_bench.py --only recsys --budget 60 > $OUT/ncu_$K.log 2>&1
    ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
  done
fi
ls -la $OUT
