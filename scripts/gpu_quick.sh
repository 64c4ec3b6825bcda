#!/bin/bash
# One B200 (under gpurun): patch-kernel parity, one bench line per headline order, optional phase profile.   usage: gpu_quick.sh TAG [prof]
tag=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest -x -q -m gpu -p no:cacheprovider tests/test_gpu_apply.py tests/test_gpu_pcg.py > gpurun_out/pytest_apply_$tag.log 2>&1
echo "apply+pcg: $(tail -1 gpurun_out/pytest_apply_$tag.log)"
BENCH="python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 20"
timeout 300 $BENCH > gpurun_out/q_p2_$tag.json 2> gpurun_out/q_p2_$tag.err
timeout 300 $BENCH --order 1 --n 255 > gpurun_out/q_p1_$tag.json 2> gpurun_out/q_p1_$tag.err
for f in gpurun_out/q_p2_$tag.json gpurun_out/q_p1_$tag.json; do python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-28s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  pcg it/s %7.1f" % (
        sys.argv[1].split("/")[-1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["pcg"]["iters_per_s"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-600:])
PY
done
if [ "$2" = "prof" ]; then bash scripts/phase_profile.sh run 2>&1 | tail -17; fi
