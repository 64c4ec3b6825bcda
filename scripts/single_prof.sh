#!/bin/bash
tag=${1:-r2c}
mkdir -p gpurun_out
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
HCURL_PATCH_SINGLE=1 timeout 200 $BENCH > gpurun_out/bench_v_single32_$tag.json 2> gpurun_out/bench_v_single32_$tag.err
HCURL_PATCH_SINGLE=1 HCURL_PATCH_FE=64 timeout 200 $BENCH > gpurun_out/bench_v_single64_$tag.json 2> gpurun_out/bench_v_single64_$tag.err
HCURL_PATCH_SINGLE=1 ncu --set full --clock-control none --import-source on -k regex:'p2s' -s 3 -c 1 -f -o gpurun_out/prof_single_$tag python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --pcg-iters 3 > gpurun_out/prof_single_$tag.log 2>&1
for f in gpurun_out/bench_v_*_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-44s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  pcg it/s %7.1f" % (
        sys.argv[1].split("bench_v_")[1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["pcg"]["iters_per_s"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
