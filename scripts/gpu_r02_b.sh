#!/bin/bash
# Round-2 call B: first run of the C minibatch loop (two streams, fused subset statistics, in-library NCCL):
# smoke, the GPU suite, the bench line, and the same bench with the scheduling options off.
TAG=${1:-r02_b}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q ${PYTEST_ARGS} > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -15 $OUT/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/bench.json 2> $OUT/bench.err
tail -c 2500 $OUT/bench.json; tail -5 $OUT/bench.err
for V in "MODL_FIT_OVERLAP=0" "MODL_FIT_GATE=0"; do
  env $V timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e > $OUT/bench_$V.json 2> $OUT/bench_$V.err
  python - <<PY
import json
d=json.load(open("$OUT/bench_$V.json"))
print("$V", "value %.0f" % d["value"], "ms/step %.4f" % d["ms_per_step"], "host %.3f" % d["host_enqueue_ms_per_step"])
PY
done
if [ -z "$SKIP_NCU" ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
      --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_list.log 2>&1
fi
ls -la $OUT
