run() { env "$@" timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 0 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'GDOF/s %.2f'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'step_ms %.4f'%d['ms_per_step'])"; }
run HCURL_PATCH_FE=64
run HCURL_PATCH_FE=64 HCURL_GATHER_CG=1
run HCURL_PATCH_FE=32
run HCURL_PATCH_FE=32 HCURL_GATHER_CG=1
run HCURL_PATCH_FE=128 HCURL_GATHER_CG=1
