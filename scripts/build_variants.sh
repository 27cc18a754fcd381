#!/bin/bash
# build tuning variants of the library (kernel experiments): build_variants/libfw_<tag>.so
set -e
cd "$(dirname "$0")/.."
mkdir -p build_variants
rm -f build_variants/*.so
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math --shared -cudart static"
S="bevy_firework_b200/csrc/fw_kernels.cu bevy_firework_b200/csrc/fw_api.cu"
build() { tag=$1; shift; nvcc $F "$@" -o build_variants/libfw_$tag.so $S; }
for spec in "$@"; do tag=${spec%%:*}; flags=${spec#*:}; build $tag $flags & done
wait
ls build_variants
