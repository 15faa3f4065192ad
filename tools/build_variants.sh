#!/bin/bash
# Development aid: build tuning variants of the library into simpleraytracing_b200/lib/variants/.
# usage: tools/build_variants.sh name "<extra nvcc/g++ -D flags>" ...
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
SRC=$ROOT/simpleraytracing_b200/csrc
OUT=$ROOT/simpleraytracing_b200/lib/variants
mkdir -p $OUT
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  tmp=$(mktemp -d)
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -ftz=false -prec-div=true -prec-sqrt=true $flags -c $SRC/kernels.cu -o $tmp/kernels.o
  g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -I$ROOT/include -I/usr/local/cuda/include $flags -c $SRC/prepack.cpp -o $tmp/prepack.o
  g++ -O2 -std=c++17 -fPIC -ffp-contract=off -fno-fast-math -I$ROOT/include -I/usr/local/cuda/include $flags -c $SRC/srt_api.cpp -o $tmp/srt_api.o
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/libsrt_b200_$name.so $tmp/kernels.o $tmp/prepack.o $tmp/srt_api.o -cudart static -L/usr/local/cuda/lib64 -lnvjpeg -Xlinker -rpath,/usr/local/cuda/lib64
  rm -rf $tmp
  echo built $name
done
