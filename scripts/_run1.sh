mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu > gpurun_out/pytest_apply.log 2>&1; tail -15 gpurun_out/pytest_apply.log
timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err; tail -c 400 gpurun_out/bench_v2.err
python -c "
import json
d=json.loads(open('gpurun_out/bench_v2.json').read().strip().splitlines()[-1])
print('GDOF/s %.2f'%d['value'], 'kernel_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'], d['pcg'])"
for fe in 32 128; do HCURL_PATCH_FE=$fe timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('FE=$fe GDOF/s %.2f'%d['value'], 'kernel_ms', d['roofline']['kernel_ms'])"; done
HCURL_PATCH_CTAS_PER_SM=1 timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('1cta/sm GDOF/s %.2f'%d['value'], 'kernel_ms', d['roofline']['kernel_ms'])"
timeout 120 python scripts/phase_profile.py 128 2>&1 | tail -12
