#!/bin/bash
# Under gpurun: one bench line per diagnostic build in build/diag/libhcurl_*.so (scripts/variant_build.sh) + the regular library.
#   usage: variant_run.sh TAG [bench args]
tag=${1:-v}; shift
mkdir -p gpurun_out
for lib in $(ls hephaestus_b200/lib/libhcurl_b200.so build/diag/libhcurl_*.so | grep -v prof); do
  name=$(basename $lib .so)
  HCURL_LIB_PATH=$PWD/$lib timeout 300 python bench.py --steps 100 --no-cpu --no-e2e --pcg-iters 10 "$@" > gpurun_out/var_${name}_$tag.json 2> gpurun_out/var_${name}_$tag.err
  python - gpurun_out/var_${name}_$tag.json <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print("%-44s GDOF/s %7.2f  frac %.3f  kernel_ms %.4f  step_ms %.4f  pcg it/s %7.1f" % (
        sys.argv[1].split("/")[-1], d["value"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["ms_per_step"], d["pcg"]["iters_per_s"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open(sys.argv[1].replace(".json", ".err")).read()[-400:])
PY
done
