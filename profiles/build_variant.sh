#!/bin/bash
# usage: build_variant.sh "<extra nvcc -D flags>"   (rebuilds libsvs_b200.so in place)
set -e
cd "$(dirname "$0")/.."
PKG="secure-video-steganography-using-ecc-and-dct_b200"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC -shared -cudart static \
     $1 -I include -I "$PKG/csrc" -o "$PKG/libsvs_b200.so" "$PKG/csrc/svs_b200.cu"
