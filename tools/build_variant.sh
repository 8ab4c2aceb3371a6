#!/bin/bash
# Build a variant of libpov_synth.so that differs from the in-tree build only in kernel_warp.cu's compile flags:
#   tools/build_variant.sh <name> [-DPOV_WARP_TMEM=0 ...]   ->  build/ab/libpov_<name>.so   (for tools/ab.sh)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
make -C parseoggvorbis_b200/csrc -j4 >/dev/null
mkdir -p build/ab
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC ${POV_VARIANT_NO_DEFAULT_FLAGS:--Xptxas -regUsageLevel=3} "$@" -c parseoggvorbis_b200/csrc/kernel_warp.cu -o build/ab/kernel_warp_$name.o
objs=$(ls build/csrc/*.o | grep -v "/kernel_warp.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -cudart static -o build/ab/libpov_$name.so $objs build/ab/kernel_warp_$name.o -lpthread
echo build/ab/libpov_$name.so
