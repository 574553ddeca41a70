#!/bin/bash
# Round-2 call D (re-entry): re-validate the state after the container was re-created: smoke, GPU suite, cycles per
# atom of the cluster dictionary kernel, loop timeline, bench line, launch list.
TAG=${1:-r02_d}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 120 python scripts/bcd_timing.py > $OUT/bcd_timing.log 2>&1; cat $OUT/bcd_timing.log
timeout 120 python scripts/loop_trace.py device 10 > $OUT/trace_device.log 2>&1; tail -40 $OUT/trace_device.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json; tail -5 $OUT/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_list.log 2>&1
ls -la $OUT
