#!/bin/bash
# Under gpurun: the warp-specialised order-2 kernel (HCURL_PATCH_STREAM=1): parity, then bench next to the default.
tag=${1:-s}
mkdir -p gpurun_out
export HCURL_PATCH_STREAM=1
timeout -s KILL 400 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_gpu_apply.py tests/test_gpu_pcg.py > gpurun_out/pytest_stream_$tag.log 2>&1
echo "stream apply+pcg: $(tail -1 gpurun_out/pytest_stream_$tag.log)"
grep -E "FAILED|Error|error" gpurun_out/pytest_stream_$tag.log | head -10
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
timeout -s KILL 200 $BENCH > gpurun_out/q_stream_$tag.json 2> gpurun_out/q_stream_$tag.err
HCURL_PATCH_STREAM=0 timeout -s KILL 200 $BENCH > gpurun_out/q_default_$tag.json 2> gpurun_out/q_default_$tag.err
for f in gpurun_out/q_stream_$tag.json gpurun_out/q_default_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  pcg it/s %7.1f" % (
        sys.argv[1].split("/")[-1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["pcg"]["iters_per_s"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-600:])
PY
done
