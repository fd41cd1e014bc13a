#!/bin/bash
# Build a tracing copy of the library (-DT4S_TRACE, gpurun_out/libt4s_trace.so) — run HERE (needs nvcc); scripts/trace_attn.py uses it on the GPU box.
set -e
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/trace_obj
for f in transformer4sed_b200/csrc/*.cu; do
  nvcc -c "$f" -o gpurun_out/trace_obj/$(basename "$f" .cu).o -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
       --expt-relaxed-constexpr -I include -I transformer4sed_b200/csrc -DT4S_TRACE &
done
wait
nvcc -shared -o scripts/probe/libt4s_trace.so gpurun_out/trace_obj/*.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -ldl
ls -la scripts/probe/libt4s_trace.so
