#!/bin/bash
# One gpurun call: GPU parity tests, the bench line (+ reference arm), the ncu launch list and
# ncu --set full captures of the dominant kernels.  Everything lands in gpurun_out/<tag>/.
#   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r01_v10'
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/pytest.log
  tail -5 $OUT/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err
  tail -c 600 $OUT/bench_ref.json
fi
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_list.log 2>&1
  for KC in ${NCU_KERNELS:-bcd_pilot:1 cd_regression:1 gemm_simt_kernel:4}; do
    K=${KC%%:*}; C=${KC##*:}
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $((4*C)) -c $C \
        -f -o $OUT/prof_$K python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_$K.log 2>&1
    ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
  done
fi
ls -la $OUT
