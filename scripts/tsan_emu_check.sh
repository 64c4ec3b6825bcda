#!/bin/bash
# ThreadSanitizer pass over the host emulation of the patch kernels written as Env programs (tests/emu/patch_emu.cpp running
# hephaestus_b200/csrc/patch_pipeline.h and patch_stream.h on one host thread per CUDA thread): a missing barrier, flag or
# wait in the kernel program shows up as a data race between the emulated threads.  No GPU needed.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/hcurl_tsan
mkdir -p $OUT
g++ -std=c++20 -O1 -g -fsanitize=thread -fno-omit-frame-pointer -fPIC -shared -pthread -I$ROOT/hephaestus_b200/csrc -I$ROOT/tests/emu \
  -I/usr/local/cuda/include $ROOT/tests/emu/patch_emu.cpp -o $OUT/libemu_tsan.so
cd $ROOT
LD_PRELOAD=$(gcc -print-file-name=libtsan.so) TSAN_OPTIONS="halt_on_error=0 report_signal_unsafe=0 exitcode=66" HCURL_EMU_LIB=$OUT/libemu_tsan.so \
  python -m pytest tests/test_patch_p1_emu.py tests/test_patch_p2_emu.py -x -q -p no:cacheprovider -k "matches_oracle or essential or zeroing_batches" 2>&1 | tee $OUT/tsan.log | tail -5
if grep -q "WARNING: ThreadSanitizer" $OUT/tsan.log; then echo "ThreadSanitizer reported races (see $OUT/tsan.log)"; exit 1; fi
echo "ThreadSanitizer: clean"
