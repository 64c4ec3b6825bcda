mkdir -p gpurun_out
( time timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 3 --no-cpu > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err ) 2>&1 | tail -4
tail -2 gpurun_out/bench_2gpu.err | cut -c1-200
python -c "
import json
d=json.loads(open('gpurun_out/bench_2gpu.json').read().strip().splitlines()[-1])
print('2gpu GDOF/s %.2f'%d['value'], 'step %.4f'%d['ms_per_step'], d['roofline']['kernel_ms'], d['pcg'], 'setup', d['setup_s'], 'launches', d['gpu_launches'])"
timeout 200 python -m pytest tests/test_gpu_multi.py -x -q -m gpu 2>&1 | tail -2
