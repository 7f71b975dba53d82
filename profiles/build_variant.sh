#!/bin/bash
# Build an A/B variant of libgd_b200.so: recompile ONE source with extra -D flags, link with the current objects.
#   profiles/build_variant.sh groupnorm.cu lib_u4.so -DGN_U=4 -DGN_MINB=4
set -e
src=$1; out=$2; shift 2
P=guided_diffusion_clip_b200
mkdir -p $P/build/ab
obj=/tmp/variant_$$.o
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -DGD_NO_FAST_MATH "$@" \
  -I include -c $P/csrc/$src -o $obj
others=$(ls $P/build/*.o | grep -v "/${src%.cu}.o")
nvcc -shared -o $P/build/ab/$out $obj $others -gencode arch=compute_100a,code=sm_100a -cudart static
rm -f $obj
echo built $P/build/ab/$out
