#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): GPU tests, the bench line, the reference arm, ncu launch list and
# full capture of the dominant kernel, phase profile, sweep over the other orders.  Outputs under gpurun_out/.
tag=${1:-r1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 400 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; tail -c 300 gpurun_out/bench_$tag.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.err; tail -c 300 gpurun_out/bench_ref_$tag.err
BENCH="python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $BENCH > gpurun_out/launches_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'hex_patch_modal' -s 3 -c 1 -f -o gpurun_out/prof_$tag $BENCH > gpurun_out/prof_$tag.log 2>&1
timeout 120 python scripts/phase_profile.py 128 > gpurun_out/phase_$tag.txt 2>&1
for cfg in "1 255" "3 85" "4 64"; do set -- $cfg; timeout 200 python bench.py --order $1 --n $2 --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p$1_$tag.json 2>gpurun_out/bench_p$1_$tag.err; done
# opt-in variants written without GPU access (round 1): parity first, then their bench lines
HCURL_TEST_PATCH_P1=1 timeout 300 python -m pytest tests/test_gpu_apply.py tests/test_gpu_pcg.py -x -q -m gpu -k "p1_patch or order1_patch" > gpurun_out/pytest_p1patch_$tag.log 2>&1; tail -3 gpurun_out/pytest_p1patch_$tag.log
for fe in 64 128 256; do HCURL_PATCH_P1=1 HCURL_PATCH_FE=$fe timeout 200 python bench.py --order 1 --n 255 --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p1patch_fe${fe}_$tag.json 2>gpurun_out/bench_p1patch_fe${fe}_$tag.err; done
HCURL_TEST_PATCH_ENV=1 timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k env_program > gpurun_out/pytest_p2env_$tag.log 2>&1; tail -3 gpurun_out/pytest_p2env_$tag.log
HCURL_PATCH_ENV=1 timeout 200 python bench.py --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2env_$tag.json 2>/dev/null
HCURL_PATCH_ENV=1 HCURL_PATCH_SPLIT=1 timeout 200 python bench.py --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2envsplit_$tag.json 2>/dev/null
HCURL_TEST_PATCH_HALF=1 timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k half_program > gpurun_out/pytest_p2half_$tag.log 2>&1; tail -3 gpurun_out/pytest_p2half_$tag.log
for hv in 1 2; do HCURL_PATCH_HALF=$hv timeout 200 python bench.py --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2half${hv}_$tag.json 2>/dev/null; done
HCURL_TEST_QAUTO=1 timeout 200 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k quad_auto > gpurun_out/pytest_qauto_$tag.log 2>&1; tail -3 gpurun_out/pytest_qauto_$tag.log
HCURL_TEST_CHEBYSHEV=1 timeout 300 python -m pytest tests/test_gpu_pcg.py -x -q -m gpu -k chebyshev > gpurun_out/pytest_cheb_$tag.log 2>&1; tail -3 gpurun_out/pytest_cheb_$tag.log
HCURL_TEST_IRIS=1 timeout 300 python -m pytest tests/test_oracle_iris_kat.py -x -q -m gpu > gpurun_out/pytest_iris_$tag.log 2>&1; tail -3 gpurun_out/pytest_iris_$tag.log
HCURL_TEST_SINGLE_TERM=1 timeout 200 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k single_term > gpurun_out/pytest_single_$tag.log 2>&1; tail -3 gpurun_out/pytest_single_$tag.log
HCURL_TEST_INTERP=1 timeout 300 python -m pytest tests/test_interp.py -x -q -m gpu > gpurun_out/pytest_interp_$tag.log 2>&1; tail -3 gpurun_out/pytest_interp_$tag.log
HCURL_TEST_SPLIT=1 timeout 300 python -m pytest tests/test_gpu_apply.py -x -q -m gpu -k split > gpurun_out/pytest_split_$tag.log 2>&1; tail -3 gpurun_out/pytest_split_$tag.log
HCURL_PATCH_SPLIT=1 timeout 200 python bench.py --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2split_$tag.json 2>/dev/null
timeout 200 python bench.py --perturbed --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2pert_$tag.json 2>/dev/null
timeout 200 python bench.py --quad mfem --steps 20 --no-cpu --no-e2e --pcg-iters 10 > gpurun_out/bench_p2mfem_$tag.json 2>/dev/null
for f in gpurun_out/bench_$tag.json gpurun_out/bench_p*_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'order', d['config']['order'], 'GDOF/s %.2f' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'kern_ms %.4f' % d['roofline']['kernel_ms'],
          'step_ms %.4f' % d['ms_per_step'], 'pcg it/s %.1f' % d['pcg']['iters_per_s'], d['roofline']['kernel'])
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat gpurun_out/bench_ref_$tag.json | tail -1 | cut -c1-400
