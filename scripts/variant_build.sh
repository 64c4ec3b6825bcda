#!/bin/bash
# Diagnostic builds of the patch kernels with other compile-time switches: recompiles hex_patch_kernels.cu only and links it
# with the objects of the regular build.    usage: variant_build.sh TAG [-DHCURL_...=v ...]   ->  build/diag/libhcurl_TAG.so
set -e
cd "$(dirname "$0")/.."
tag=$1; shift
mkdir -p build/diag
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -x cu "$@" -c hephaestus_b200/csrc/hex_patch_kernels.cu -o build/diag/hpk_$tag.o
objs=$(ls build/obj/*.o | grep -v hex_patch_kernels.o)
nvcc -shared -o build/diag/libhcurl_$tag.so $objs build/diag/hpk_$tag.o -gencode arch=compute_100a,code=sm_100a -lcudart -ldl
n=$(cuobjdump -sass -fun '_ZN5hcurl19hex_patch_p2_kernelILi32ELb0ELb0ELb0EEEvNS_9PatchArgsE' build/diag/libhcurl_$tag.so 2>/dev/null | grep -cE "^\s+/\*[0-9a-f]{4,5}\*/")
r=$(cuobjdump -res-usage build/diag/libhcurl_$tag.so 2>/dev/null | grep -A1 'hex_patch_p2_kernelILi32ELb0ELb0ELb0' | grep -o "REG:[0-9]* STACK:[0-9]*")
echo "$tag: $n SASS instructions, $r"
