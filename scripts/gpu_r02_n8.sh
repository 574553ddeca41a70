#!/bin/bash
# Round-2 eight-GPU call: the bench line at N = 8 (strong scaling primary, weak alongside, parity, Amdahl).
TAG=${1:-r02_n8}
N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export PYTHONUNBUFFERED=1
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 20 --warmup 5 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
tail -3 $OUT/bench_n$N.err
python - <<PY
import json
d=json.load(open("$OUT/bench_n$N.json"))
print("N=%d %s: value %.0f ms/step %.4f host %.3f | e2e %s" % (d["n_gpus"], d["scaling"], d["value"], d["ms_per_step"], d["host_enqueue_ms_per_step"], d["e2e"] and ("%.0f (%.4f ms)" % (d["e2e"]["value"], d["e2e"]["ms_per_step"]))))
for k in ("weak", "strong"):
    if k in d: print(k, "value %.0f ms/step %.4f host %.3f" % (d[k]["value"], d[k]["ms_per_step"], d[k]["host_enqueue_ms_per_step"]))
print("parity", d["parity"]); print("amdahl", d.get("amdahl"))
print({k: round(v["ms"]*1e3,1) for k,v in d["roofline"]["phases"].items()})
PY
ls $OUT
