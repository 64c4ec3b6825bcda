mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_final.log 2>&1; tail -2 gpurun_out/pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 200 gpurun_out/bench_final.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1])
print('GDOF/s %.2f'%d['value'], 'step %.4f'%d['ms_per_step'], 'kernel %.4f'%d['roofline']['kernel_ms'], 'frac %.3f'%d['roofline']['frac'], 'traffic', d['roofline']['traffic'], 'pcg', d['pcg']['iters_per_s'], 'e2e', d['e2e']['value'], 'cpu', d['cpu_baseline']['value'], d['clocks'])"
