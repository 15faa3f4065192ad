#!/bin/bash
# usage: tools/run_variants.sh name1 name2 ...   (development aid; run on the GPU box)
mkdir -p gpurun_out
for v in "$@"; do
  SRT_LIB=simpleraytracing_b200/lib/variants/libsrt_b200_$v.so timeout 120 python tools/variant_bench.py 2>&1 | tail -1 | tee -a gpurun_out/variants.log
done
