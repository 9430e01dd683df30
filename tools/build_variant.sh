#!/usr/bin/env bash
# tools/build_variant.sh <name> <nvcc defines...> -- an experimental copy of the library with extra -D flags, for A/B runs on the
# GPU box (select it with CUMF_ALS_LIB=cumf_als_b200/libcumf_als_b200_<name>.so); never loaded by default.
set -euo pipefail
cd "$(dirname "$0")/.."
name=$1; shift
obj=build/obj_$name
mkdir -p $obj
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default --expt-relaxed-constexpr -Wno-deprecated-gpu-targets"
pids=()
for src in als_api.cu gram_simt.cu gram_tc.cu gram_tc2.cu gram_tc2_a.cu gram_tc2_b.cu gram_tc2_c.cu cg.cu rmse.cu synth.cu host_io.cpp; do
  nvcc $FLAGS "$@" -c cumf_als_b200/csrc/$src -o $obj/$src.o &
  pids+=($!)
done
for p in "${pids[@]}"; do wait $p; done
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o cumf_als_b200/libcumf_als_b200_$name.so $obj/*.o -lcublas -Xlinker -rpath=/usr/local/cuda/lib64
ls -la cumf_als_b200/libcumf_als_b200_$name.so
