#!/bin/bash
# build an alternative library with extra nvcc flags into footprint-tools_b200/lib_alt/<name>/libfpt_b200.so
name=$1; shift
cd /root/repo/footprint-tools_b200
mkdir -p lib_alt/$name
for f in fpt_score fpt_fast fpt_fused fpt_warp fpt_ops fpt_fdr fpt_text fpt_ingest fpt_segment fpt_api; do
  nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden --expt-relaxed-constexpr "$@" -c csrc/$f.cu -o lib_alt/$name/$f.o &
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o lib_alt/$name/libfpt_b200.so lib_alt/$name/*.o -lcudart_static -lpthread -ldl -lrt
rm -f lib_alt/$name/*.o
ls -la lib_alt/$name/
