#!/bin/bash
# First GPU call of a round (run under gpurun, one B200): parity of every headline-kernel variant written without GPU
# access, then one bench line each, so that the default can be chosen from measurements.  ~5 minutes.  Outputs under
# gpurun_out/; the summary table is printed last.
tag=${1:-r2a}
mkdir -p gpurun_out
run_tests() {  # name, env switch, pytest args...
  local name=$1 sw=$2; shift 2
  env $sw=1 timeout 400 python -m pytest -x -q -m gpu -p no:cacheprovider "$@" > gpurun_out/pytest_${name}_$tag.log 2>&1
  echo "$name: $(tail -1 gpurun_out/pytest_${name}_$tag.log)"
}
run_tests default HCURL_NONE tests/test_gpu_apply.py tests/test_gpu_pcg.py tests/test_zy_fullsize_properties.py
run_tests split HCURL_TEST_SPLIT tests/test_gpu_apply.py -k split
run_tests env HCURL_TEST_PATCH_ENV tests/test_gpu_apply.py -k env_program
run_tests half HCURL_TEST_PATCH_HALF tests/test_gpu_apply.py -k half_program
run_tests p1patch HCURL_TEST_PATCH_P1 tests/test_gpu_apply.py tests/test_gpu_pcg.py -k "p1_patch or order1_patch"
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
bench() {  # name, environment assignments...
  local name=$1; shift
  env "$@" timeout 200 $BENCH > gpurun_out/bench_v_${name}_$tag.json 2> gpurun_out/bench_v_${name}_$tag.err
}
bench default HCURL_NONE=0
bench split HCURL_PATCH_SPLIT=1
bench env HCURL_PATCH_ENV=1
bench envsplit HCURL_PATCH_ENV=1 HCURL_PATCH_SPLIT=1
bench half96 HCURL_PATCH_HALF=1
bench half80 HCURL_PATCH_HALF=2
for c in 2 3; do bench half80_cta$c HCURL_PATCH_HALF=2 HCURL_PATCH_CTAS_PER_SM=$c; done
HCURL_PATCH_P1=1 timeout 200 python bench.py --order 1 --n 255 --steps 50 --no-cpu --no-e2e --pcg-iters 20 > gpurun_out/bench_v_p1patch_$tag.json 2> gpurun_out/bench_v_p1patch_$tag.err
timeout 200 python bench.py --order 1 --n 255 --steps 50 --no-cpu --no-e2e --pcg-iters 20 > gpurun_out/bench_v_p1generic_$tag.json 2> gpurun_out/bench_v_p1generic_$tag.err
for f in gpurun_out/bench_v_*_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-44s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  pcg it/s %7.1f  %s" % (
        sys.argv[1].split("bench_v_")[1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"],
        d["pcg"]["iters_per_s"], d["roofline"]["kernel"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
