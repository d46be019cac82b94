#!/bin/bash
# A/B builds of the library for same-box timing:  profiles/ab_build.sh NAME "-DKNOB=..."  -> ab/libxrsd_NAME.so
# (ab/ is git-ignored but travels to the GPU box; run a build through XRSD_B200_LIB=ab/libxrsd_NAME.so)
set -e
NAME=$1; shift
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=$ROOT/ab/_build_$NAME
mkdir -p $OUT
for f in abi intensity_kernel table_kernel fit_kernels feature_kernel; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" \
     -c $ROOT/xrsdkit_b200/csrc/$f.cu -o $OUT/$f.o > $OUT/$f.log 2>&1 &
done
wait
/usr/local/cuda/bin/nvcc -shared -o $ROOT/ab/libxrsd_$NAME.so $OUT/*.o -lcudart
echo $ROOT/ab/libxrsd_$NAME.so
