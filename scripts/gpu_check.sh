#!/bin/bash
# One B200 (run under gpurun): GPU parity tests of the patch kernels, then bench lines of the headline configurations.
# Outputs under gpurun_out/; the summary table is printed last.    usage: gpu_check.sh TAG [full]
tag=${1:-r2}
mkdir -p gpurun_out
timeout 900 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_gpu_apply.py tests/test_gpu_pcg.py > gpurun_out/pytest_apply_$tag.log 2>&1
echo "apply+pcg: $(tail -1 gpurun_out/pytest_apply_$tag.log)"
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
bench() { local name=$1; shift; env "$@" timeout 300 $BENCH > gpurun_out/bench_v_${name}_$tag.json 2> gpurun_out/bench_v_${name}_$tag.err; }
bench p2_fe32 HCURL_X=0
bench p2_fe64 HCURL_PATCH_FE=64
HCURL_X=0 timeout 300 $BENCH --order 1 --n 255 > gpurun_out/bench_v_p1_fe128_$tag.json 2> gpurun_out/bench_v_p1_fe128_$tag.err
if [ "$2" = "full" ]; then
  for c in 4 5; do bench p2_fe32_cta$c HCURL_PATCH_CTAS_PER_SM=$c; done
  HCURL_PATCH_FE=64 timeout 300 $BENCH --order 1 --n 255 > gpurun_out/bench_v_p1_fe64_$tag.json 2> gpurun_out/bench_v_p1_fe64_$tag.err
  HCURL_PATCH_FE=256 timeout 300 $BENCH --order 1 --n 255 > gpurun_out/bench_v_p1_fe256_$tag.json 2> gpurun_out/bench_v_p1_fe256_$tag.err
fi
ncu --set full --clock-control none --import-source on -k regex:'hex_patch_p2' -s 3 -c 1 -f -o gpurun_out/prof_p2_$tag python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3 > gpurun_out/prof_p2_$tag.log 2>&1
for f in gpurun_out/bench_v_*_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-36s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  share %.3f  pcg it/s %7.1f" % (
        sys.argv[1].split("bench_v_")[1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"],
        d["roofline"]["kernel_share_of_step"], d["pcg"]["iters_per_s"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-400:])
PY
done
