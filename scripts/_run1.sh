mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_apply.py tests/test_gpu_pcg.py tests/test_reference_meshes.py -x -q -m gpu 2>&1 | tail -2
run() { env "$@" timeout 100 python bench.py --steps 30 --warmup 5 --no-cpu --no-e2e --pcg-iters 20 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$*', 'GDOF/s %.2f'%d['value'], 'kernel_ms %.4f'%d['roofline']['kernel_ms'], 'step_ms %.4f'%d['ms_per_step'], 'pcg ms/it %.4f'%d['pcg']['ms_per_iter'], 'setup %.1f'%d['setup_s'])"; }
run HCURL_PATCH_FE=32
run HCURL_PATCH_FE=64
