mkdir -p gpurun_out/r01_n2
timeout 600 python -m pytest tests/test_gpu_distributed.py -m gpu -x -q 2>&1 | tail -4
for sc in weak strong; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scaling $sc > gpurun_out/r01_n2/bench_$sc.json 2> gpurun_out/r01_n2/bench_$sc.err
python -c "
import json
d = json.load(open('gpurun_out/r01_n2/bench_$sc.json'))
print('$sc', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print({k: round(v['ms']*1e3,1) for k, v in d['roofline']['phases'].items()})
" || tail -5 gpurun_out/r01_n2/bench_$sc.err
done
CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 20 --warmup 3 --no-cpu > gpurun_out/r01_n2/bench_n1.json; python -c "
import json
d = json.load(open('gpurun_out/r01_n2/bench_n1.json'))
print('n1', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print({k: round(v['ms']*1e3,1) for k, v in d['roofline']['phases'].items()})
"
