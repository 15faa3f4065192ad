mkdir -p gpurun_out
for v in c8t256 c6t256 c5t256 c4t512; do
  echo "== $v" >> gpurun_out/variants.log
  SRT_LIB=simpleraytracing_b200/lib/variants/libsrt_b200_$v.so SRT_ONLY_FILTERED=1 SRT_PPTS=4,8 timeout 120 python tools/perf_sweep.py 1024 2048 >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log | cut -c1-200
