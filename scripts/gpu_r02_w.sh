#!/bin/bash
# Round-2 call W: graph-replayed step (capture + cudaGraphExecUpdate per call): parity test, suite, timelines, bench.
TAG=${1:-r02_w}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
timeout 300 python -m pytest tests/test_gpu_dict_fact.py -m gpu -x -q -k "graph_replay or schedules_agree or reproducible" > $OUT/pytest_graph.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_graph.log; tail -15 $OUT/pytest_graph.log
timeout 120 python scripts/loop_trace.py pinned 16 > $OUT/trace_pinned.log 2>&1; tail -6 $OUT/trace_pinned.log | cut -c1-260
timeout 120 python scripts/loop_trace.py device 8 > $OUT/trace_device.log 2>&1; tail -3 $OUT/trace_device.log | cut -c1-260
MODL_FIT_GRAPH=1 timeout 120 python scripts/loop_trace.py pinned 16 > $OUT/trace_pinned_nograph.log 2>&1; tail -3 $OUT/trace_pinned_nograph.log | cut -c1-260
timeout 600 python bench.py --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -3 $OUT/bench.err
MODL_FIT_GRAPH=1 timeout 600 python bench.py --no-cpu > $OUT/bench_nograph.json 2> $OUT/bench_nograph.err
python - <<PY
import json
for name in ("bench", "bench_nograph"):
    try:
        d=json.load(open("$OUT/%s.json" % name))
        print(name, "value %.0f ms/step %.4f (min %.4f max %.4f) host %.3f  e2e %.0f (%.4f ms, host %.3f)" % (d["value"], d["ms_per_step"], d["run"]["ms_per_step_min"], d["run"]["ms_per_step_max"], d["host_enqueue_ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["host_enqueue_ms_per_step"]))
    except Exception as e:
        print(name, "failed", e)
PY
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $OUT/pytest_gpu.log; tail -4 $OUT/pytest_gpu.log
