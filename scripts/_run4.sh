mkdir -p gpurun_out
( time timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 30 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_4gpu.json 2> gpurun_out/bench_4gpu.err ) 2>&1 | tail -4
tail -2 gpurun_out/bench_4gpu.err | cut -c1-200
python -c "
import json
d=json.loads(open('gpurun_out/bench_4gpu.json').read().strip().splitlines()[-1])
print('4gpu GDOF/s %.2f'%d['value'], 'step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms'], d['pcg'], 'setup', d['setup_s'], 'launches', d['gpu_launches'])"
