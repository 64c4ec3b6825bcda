#!/bin/bash
# AddressSanitizer + UBSan pass over the host-only sources (no GPU): builds them with a stub for the error string and
# runs scripts/asan_host_drive.py against that build.
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=/tmp/hcurl_asan
mkdir -p $OUT
cat > $OUT/stub.cpp <<'EOS'
#include <cstdarg>
#include <cstdio>
#include <string>
namespace hcurl {
static std::string g_err;
void set_error(const char *fmt, ...) { char buf[1024]; va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap); g_err = buf; }
}
extern "C" const char *hcurl_last_error(void) { return hcurl::g_err.c_str(); }
EOS
g++ -std=c++17 -O1 -g -fsanitize=address,undefined -fno-omit-frame-pointer -shared -fPIC -I/usr/local/cuda/include -pthread \
  $ROOT/hephaestus_b200/csrc/patch.cpp $ROOT/hephaestus_b200/csrc/partition.cpp $ROOT/hephaestus_b200/csrc/ndspace.cpp $ROOT/hephaestus_b200/csrc/exodus.cpp \
  $ROOT/hephaestus_b200/csrc/basis.cpp $OUT/stub.cpp -o $OUT/libhost_asan.so
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0 HCURL_ASAN_LIB=$OUT/libhost_asan.so python $ROOT/scripts/asan_host_drive.py
