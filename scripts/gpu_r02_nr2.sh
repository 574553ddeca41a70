#!/bin/bash
# after "plain launches while the dictionary update is grid-wide": next-row configs at the default settings + the GPU suite
mkdir -p gpurun_out/r02_nr2
export PYTHONUNBUFFERED=1
timeout 400 python scripts/next_rows_bench.py --budget 200 --out gpurun_out/r02_nr2/next_rows.json > gpurun_out/r02_nr2/next_rows.log 2>&1
python - <<PY
import json
d=json.load(open("gpurun_out/r02_nr2/next_rows.json"))
for r in d:
    if "value_partial_fit_only" in r:
        print(r["row"], "whole %.0f/s  partial_fit only %.0f/s  dev_s %.3f" % (r["value"], r["value_partial_fit_only"], r["partial_fit_device_seconds"]), r.get("cuda_graphs"), "cpu", r.get("cpu_reference", {}).get("value"))
    else:
        print(r["row"], {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else None) for k, v in r.items() if k.startswith("batch") or k == "cpu_reference"})
PY
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02_nr2/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02_nr2/pytest_gpu.log; tail -4 gpurun_out/r02_nr2/pytest_gpu.log
