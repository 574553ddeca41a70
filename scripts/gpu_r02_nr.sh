mkdir -p gpurun_out/r02_nr
run() { # name env...
  name=$1; shift
  env "$@" timeout 200 python scripts/next_rows_bench.py --only $ONLY --budget 100 --out gpurun_out/r02_nr/$name.json > gpurun_out/r02_nr/$name.log 2>&1
  python - <<PY
import json
d=json.load(open("gpurun_out/r02_nr/$name.json"))
for r in d:
    if "value_partial_fit_only" in r:
        print("$name", r["row"], "whole %.0f/s  partial_fit only %.0f/s  dev_s %.3f" % (r["value"], r["value_partial_fit_only"], r["partial_fit_device_seconds"]), r.get("cuda_graphs"))
PY
}
ONLY=fmri run fmri_g1_nogate MODL_FIT_GRAPH=1 MODL_FIT_GATE=0
ONLY=fmri run fmri_g0_nogate MODL_FIT_GRAPH=0 MODL_FIT_GATE=0
ONLY=image run image_g0 MODL_FIT_GRAPH=0
ONLY=image run image_g1 MODL_FIT_GRAPH=1
ONLY=image run image_g2 MODL_FIT_GRAPH=2
