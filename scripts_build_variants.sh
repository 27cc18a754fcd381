#!/bin/bash
# build tuning variants of the library (kernel experiments): build_variants/libfw_<tag>.so
set -e
cd "$(dirname "$0")"
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -prec-div=true -prec-sqrt=true -ftz=false -Xcompiler -fPIC,-O2,-ffp-contract=off,-fno-fast-math --shared -cudart static"
S="bevy_firework_b200/csrc/fw_kernels.cu bevy_firework_b200/csrc/fw_api.cu"
build() { tag=$1; shift; nvcc $F "$@" -Xptxas -v -o build_variants/libfw_$tag.so $S 2>&1 | grep -A1 "update_kernelILb0ELb0" | grep Used | sed "s/^/$tag: /"; }
build base &
build minb5 -DFW_MINB=5 &
build minb6 -DFW_MINB=6 &
build cs -DFW_CS=1 &
wait
build cs_minb5 -DFW_CS=1 -DFW_MINB=5 &
build t128 -DFW_TILE=128 &
build t512 -DFW_TILE=512 &
build t128_minb10 -DFW_TILE=128 -DFW_MINB=10 &
wait
