run() { env "$@" timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'GDOF/s %.2f'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'step_ms %.4f'%d['ms_per_step'], 'pcg ms/it %.4f'%d['pcg']['ms_per_iter'], 'traffic', d['roofline']['traffic'])"; }
run HCURL_PATCH_FLAGS=0
run HCURL_PATCH_FLAGS=1
run HCURL_PATCH_FLAGS=2
run HCURL_PATCH_FLAGS=3
HCURL_PATCH_FLAGS=3 timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu 2>&1 | tail -2
