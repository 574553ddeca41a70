#!/bin/bash
# Round-2 confirmation call: smoke, the whole GPU suite (drift numbers printed), bench line + reference arm, ncu launch list,
# ncu --set full of the dominant kernels, BASELINE configs[4] at its stated size.
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" >> $OUT/smoke.log; tail -2 $OUT/smoke.log
timeout 1200 python -m pytest tests -m gpu -q -s -rs > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log
grep -i "drift over\|passed\|failed\|error" $OUT/pytest_gpu.log | tail -6
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
python - <<PY
import json
d=json.load(open("$OUT/bench.json"))
print("bench value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
print("cpu", d["cpu_baseline"]["value"], "parity", d["parity"], "dense", d["extra"]["dense_codes"]["ms_per_step"], "launches", d["gpu_launches"])
PY
tail -3 $OUT/bench.err
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; tail -c 400 $OUT/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv \
    --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_list.log 2>&1
for KC in bcd_blocked:1:4 cd_regression:1:4 tc_gemm_kernel:3:9; do
  K=$(echo $KC | cut -d: -f1); C=$(echo $KC | cut -d: -f2); S=$(echo $KC | cut -d: -f3)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s $S -c $C \
      -f -o $OUT/prof_$K python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > $OUT/ncu_$K.log 2>&1
  ncu -i $OUT/prof_$K.ncu-rep --page raw --csv > $OUT/prof_$K.raw.csv 2>/dev/null
done
if [ -z "$SKIP_RECSYS" ]; then
  timeout 900 python scripts/recsys_scale.py --rows 1000000 --out $OUT/recsys_scale.json > $OUT/recsys_scale.log 2>&1; tail -c 1200 $OUT/recsys_scale.log
fi
ls -la $OUT
