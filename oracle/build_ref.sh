#!/usr/bin/env bash
# oracle/build_ref.sh -- TEST INFRASTRUCTURE, not product code.
#
# Compiles the UNMODIFIED reference sources, where they lie under
# /root/reference (read-only), for sm_100a with the two-symbol shim in
# oracle/ref_shim/.  Outputs go ONLY to oracle/_ref/ (git-ignored, but shipped
# to the GPU box by gpurun):
#   libref_als_cg.so  reference as shipped (#define USE_CG, CG_ITER 6)   als.cu:28,32
#   libref_als_lu.so  same sources with the single line als.cu:28 disabled
#                     (done on a throw-away copy in a mktemp dir, never stored)
#   ref_main_cg / ref_main_lu   the reference CLI (main.cpp) on top of each
# No reference source is copied into the repository.
set -euo pipefail
REF=${CUMF_REFERENCE_DIR:-/root/reference}
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/_ref"
if [ ! -d "$REF" ]; then
  echo "build_ref: $REF not present (GPU box?) -- keeping prebuilt files in $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT
NVCC=${NVCC:-nvcc}
FLAGS=(-O3 -std=c++14 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a
       -include "$HERE/ref_shim/shim.h" -I"$REF" -w -DDEBUG)

obj() { # obj <src> <out.o> [extra flags]
  local src=$1 out=$2; shift 2
  "$NVCC" "${FLAGS[@]}" "$@" -c "$src" -o "$out"
}
obj "$REF/cg.cu"               "$TMP/cg.o" &
obj "$REF/device_utilities.cu" "$TMP/du.o" &
obj "$REF/host_utilities.cpp"  "$TMP/hu.o" &
obj "$REF/main.cpp"            "$TMP/main.o" &
"$NVCC" -O3 -std=c++14 -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a -w \
    -c "$HERE/ref_shim/csrmm2_shim.cu" -o "$TMP/shim.o" &
# als.cu is compiled through ref_shim/ref_hooks.cu, which #includes it verbatim from where it
# lies and appends extern "C" launchers for the stage-level golden vectors.
obj "$HERE/ref_shim/ref_hooks.cu" "$TMP/als_cg.o" -DREF_ALS_SOURCE="\"$REF/als.cu\"" &
# LU oracle: als.cu with its line-28 `#define USE_CG` commented out, nothing else.
sed 's|^#define USE_CG|//#define USE_CG|' "$REF/als.cu" > "$TMP/als_lu.cu"
obj "$HERE/ref_shim/ref_hooks.cu" "$TMP/als_lu.o" -DREF_ALS_SOURCE="\"$TMP/als_lu.cu\"" &
wait
LIBS=(-lcublas -lcusparse)
for v in cg lu; do
  "$NVCC" -shared -o "$OUT/libref_als_$v.so" "$TMP/als_$v.o" "$TMP/cg.o" "$TMP/du.o" "$TMP/hu.o" \
      "$TMP/shim.o" "${LIBS[@]}"
  "$NVCC" -o "$OUT/ref_main_$v" "$TMP/main.o" "$TMP/als_$v.o" "$TMP/cg.o" "$TMP/du.o" "$TMP/hu.o" \
      "$TMP/shim.o" "${LIBS[@]}"
done
ls -la "$OUT"
