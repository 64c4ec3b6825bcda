#!/bin/bash
# Phase profile of the patch pipeline (diagnostic build, never the measured one): rebuilds the library with
# -DHCURL_PHASE_PROFILE into build/diag/ (run HERE before gpurun: `bash scripts/phase_profile.sh build [extra nvcc flags]`),
# then on the GPU box `bash scripts/phase_profile.sh run [bench args]` prints cycles per phase per pipeline iteration.
set -e
cd "$(dirname "$0")/.."
OUT=build/diag/libhcurl_b200_prof.so
if [ "$1" = "build" ]; then
  shift
  mkdir -p build/diag/obj
  FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -x cu -DHCURL_PHASE_PROFILE $*"
  pids=()
  for f in basis.cpp api.cu comm.cu tet_apply.cu hex_fast.cu hex_patch.cu hex_patch_kernels.cu patch.cpp ndspace.cpp partition.cpp exodus.cpp debug.cu interp.cu; do
    nvcc $FLAGS -c hephaestus_b200/csrc/$f -o build/diag/obj/${f%.*}.o & pids+=($!)
  done
  nvcc $FLAGS -c hephaestus_b200/csrc/hex_apply.cu -o build/diag/obj/hex_dispatch.o & pids+=($!)
  for p in 1 2 3 4; do nvcc $FLAGS -DHC_P=$p -c hephaestus_b200/csrc/hex_apply.cu -o build/diag/obj/hex_apply_p$p.o & pids+=($!); done
  for p in "${pids[@]}"; do wait $p; done
  nvcc -shared -o $OUT build/diag/obj/*.o -gencode arch=compute_100a,code=sm_100a -lcudart -ldl
  echo built $OUT
  exit 0
fi
shift || true
HCURL_LIB_PATH=$PWD/$OUT python scripts/phase_profile.py "$@"
