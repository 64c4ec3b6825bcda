#!/bin/bash
# ThreadSanitizer pass over the host emulation of the patch kernels written as Env programs (tests/emu/patch_p1_emu.cpp,
# patch_p2_emu.cpp running hephaestus_b200/csrc/hex_patch_p1_body.h / hex_patch_p2_body.h on one host thread per CUDA thread): a missing CTA barrier in the kernel
# program shows up as a data race between the emulated threads.  No GPU needed.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/hcurl_tsan
mkdir -p $OUT
for k in p1 p2; do
  g++ -std=c++20 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -fPIC -shared -pthread -I$ROOT/hephaestus_b200/csrc -I$ROOT/tests/emu \
    -I/usr/local/cuda/include $ROOT/tests/emu/patch_${k}_emu.cpp -o $OUT/lib${k}emu_tsan.so
done
cd $ROOT
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66" HCURL_EMU_LIB=$OUT/libp1emu_tsan.so HCURL_EMU2_LIB=$OUT/libp2emu_tsan.so \
  python -m pytest tests/test_patch_p1_emu.py tests/test_patch_p2_emu.py -x -q -k "matches_oracle or essential" 2>&1 | tee $OUT/tsan.log | tail -5
if grep -q "WARNING: ThreadSanitizer" $OUT/tsan.log; then echo "ThreadSanitizer reported races (see $OUT/tsan.log)"; exit 1; fi
echo "ThreadSanitizer: clean"
