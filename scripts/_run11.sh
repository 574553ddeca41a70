mkdir -p gpurun_out/r01_n2
for ov in 0 1; do
MODL_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/r01_n2/bench_ov$ov.json 2> gpurun_out/r01_n2/bench_ov$ov.err
python -c "
import json
d = json.load(open('gpurun_out/r01_n2/bench_ov$ov.json'))
print('overlap $ov', round(d['value']), d['ms_per_step'], 'host', d['host_enqueue_ms_per_step'], 'e2e', round(d['e2e']['value']), d['e2e']['ms_per_step'])
"
done
