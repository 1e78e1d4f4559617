#!/bin/bash
# build a compile-time variant of the library: tools/variant.sh name -DLIP_X=... ...  -> build/name.so (prints registers / spills of the ICP kernels)
name=$1; shift
mkdir -p build
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC -DLIP_ARCH=100 "$@" -Xptxas -v \
  -o build/$name.so 4d_plant_analysis_b200/csrc/local_icp.cu 2>&1 | grep -A3 "icp_group_kernel\|icp_big_kernel\|blend_tile_kernel" | grep "Used\|spill" | paste - - | sed "s/^/$name: /"
