#!/usr/bin/env bash
# A/B and instrumented builds: build_variant.sh NAME [extra nvcc flags...] -> cdv-slam_b200/lib/libpgba_NAME.so
# (git-ignored, shipped to the GPU box; select with PGBA_LIB=cdv-slam_b200/lib/libpgba_NAME.so).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
NAME="$1"; shift
OUT="$HERE/../lib"
OBJ="$OUT/obj_$NAME"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v $*"
SRCS="ba_plan ba_numeric ba_bigsolve ba_bignd ba_api ba_neighbors corr_kernels corr_tma pgo"
pids=()
for f in $SRCS; do
  "$NVCC" $FLAGS -c "$HERE/$f.cu" -o "$OBJ/$f.o" > "$OBJ/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$OBJ"/*.log; exit 1; fi
objs=""
for f in $SRCS; do objs="$objs $OBJ/$f.o"; done
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libpgba_$NAME.so" $objs -lcudart
ls -la "$OUT/libpgba_$NAME.so"
