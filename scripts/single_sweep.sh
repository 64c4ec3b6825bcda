#!/bin/bash
# one-thread-per-element order-2 patch kernel: parity, then bench lines at FE = 32 / 64 and CTA caps
tag=${1:-r2b}
mkdir -p gpurun_out
timeout 300 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_gpu_apply.py -k single_program > gpurun_out/pytest_single_$tag.log 2>&1
echo "single: $(tail -1 gpurun_out/pytest_single_$tag.log)"
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
bench() { local name=$1; shift; env "$@" timeout 200 $BENCH > gpurun_out/bench_v_${name}_$tag.json 2> gpurun_out/bench_v_${name}_$tag.err; }
bench single32 HCURL_PATCH_SINGLE=1
bench single64 HCURL_PATCH_SINGLE=1 HCURL_PATCH_FE=64
for c in 3 4 5; do bench single32_cta$c HCURL_PATCH_SINGLE=1 HCURL_PATCH_CTAS_PER_SM=$c; done
bench single64_cta2 HCURL_PATCH_SINGLE=1 HCURL_PATCH_FE=64 HCURL_PATCH_CTAS_PER_SM=2
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
