mkdir -p gpurun_out
python bench.py --steps 50 --warmup 5 > gpurun_out/bench_p2.json 2> gpurun_out/bench_p2.err; tail -c 600 gpurun_out/bench_p2.err
for cfg in "1 255" "3 85" "4 64"; do set -- $cfg; python bench.py --order $1 --n $2 --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p$1.json 2>gpurun_out/bench_p$1.err; done
python bench.py --perturbed --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2_pert.json 2>&1
python bench.py --quad mfem --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2_mfem.json 2>&1
for f in gpurun_out/bench_p*.json; do echo $f; python -c "
import sys,json
d=json.loads(open('$f').read().strip().splitlines()[-1])
print(d['config']['order'], 'GDOF/s %.2f'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'kern_ms %.3f'%d['roofline']['kernel_ms'], 'step_ms %.3f'%d['ms_per_step'], 'pcg it/s %.1f'%d['pcg']['iters_per_s'], 'fp64 %.1f'%d['roofline']['fp64_fma_peak_tflops_measured'], d['clocks'], d.get('e2e'), d.get('cpu_baseline'), 'setup', d['setup_s'])
"; done
