#!/bin/bash
# Round evidence on one B200 (run under gpurun): GPU tests, the bench line, the reference arm, ncu launch list and full
# capture of the dominant kernel, the other orders / kernels.  Outputs under gpurun_out/.     usage: final_run.sh TAG [notests]
tag=${1:-r2}
mkdir -p gpurun_out
if [ "$2" != "notests" ]; then timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log; fi
timeout 500 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 300 gpurun_out/bench_$tag.err
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 300 gpurun_out/bench_ref_$tag.err
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $BENCH > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hex_patch_p2' -s 3 -c 1 -f -o gpurun_out/prof_$tag $BENCH > gpurun_out/prof_$tag.log 2>&1
S="--steps 50 --no-cpu --no-e2e --pcg-iters 20"
for cfg in "1 255" "3 85" "4 64"; do set -- $cfg; timeout 300 python bench.py --order $1 --n $2 $S > gpurun_out/bench_p$1_$tag.json 2>gpurun_out/bench_p$1_$tag.err; done
timeout 300 python bench.py --perturbed $S > gpurun_out/bench_p2pert_$tag.json 2>/dev/null
timeout 300 python bench.py --quad mfem $S > gpurun_out/bench_p2mfem_$tag.json 2>/dev/null
HCURL_PATCH_STREAM=0 timeout 300 python bench.py $S > gpurun_out/bench_p2pipeline32_$tag.json 2>/dev/null
HCURL_PATCH_FE=64 timeout 300 python bench.py $S > gpurun_out/bench_p2stream64_$tag.json 2>/dev/null
timeout 300 python bench.py --geom tet --order 2 --n 64 $S > gpurun_out/bench_ptet2_$tag.json 2>/dev/null
timeout 300 python bench.py --geom tet --order 1 --n 100 $S > gpurun_out/bench_ptet1_$tag.json 2>/dev/null
for f in gpurun_out/bench_$tag.json gpurun_out/bench_p*_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'order', d['config']['order'], 'GDOF/s %.2f' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'kern_ms %.4f' % d['roofline']['kernel_ms'],
          'step_ms %.4f' % d['ms_per_step'], 'pcg it/s %.1f' % d['pcg']['iters_per_s'], d['roofline']['kernel'][:60])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat gpurun_out/bench_ref_$tag.json | tail -1 | cut -c1-400
