#!/bin/bash
# Builds a tuning variant of libabip_gpu.so into build/variants/libabip_gpu_<name>.so (git-ignored, ships with gpurun).
# usage: tools/build_variant.sh <name> "<extra nvcc flags, e.g. -DABIP_PHASE_TIMING -DABIP_CHUNK=192>"
# select it at run time with ABIP_GPU_LIB=build/variants/libabip_gpu_<name>.so
set -e
cd "$(dirname "$0")/.."
name=$1; flags=$2
out=build/variants/$name; mkdir -p $out
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ccbin /usr/bin/g++ $flags"
src=abip_b200/csrc
$NV -c $src/lp_engine.cu -o $out/lp_engine.o &
$NV -c $src/qcp_engine.cu -o $out/qcp_engine.o &
g++ -O2 -std=c++17 -fPIC $flags -c $src/lp_host.cpp -o $out/lp_host.o &
g++ -O2 -std=c++17 -fPIC $flags -c $src/qcp_host.cpp -o $out/qcp_host.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -ccbin /usr/bin/g++ -o build/variants/libabip_gpu_$name.so $out/*.o -cudart static
rm -rf $out
echo built build/variants/libabip_gpu_$name.so
