#!/usr/bin/env bash
# Build libpgba.so (sm_100a only) in-tree: cdv-slam_b200/lib/libpgba.so.  No torch headers are involved.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT/obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC -Xptxas -v"
pids=()
for f in ba_plan ba_numeric ba_bigsolve ba_bignd ba_api ba_neighbors corr_kernels corr_tma pgo; do
  "$NVCC" $FLAGS -c "$HERE/$f.cu" -o "$OUT/obj/$f.o" > "$OUT/obj/$f.log" 2>&1 &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "$OUT"/obj/*.log; exit 1; fi
"$NVCC" -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libpgba.so" \
    "$OUT/obj/ba_plan.o" "$OUT/obj/ba_numeric.o" "$OUT/obj/ba_bigsolve.o" "$OUT/obj/ba_bignd.o" "$OUT/obj/ba_api.o" "$OUT/obj/ba_neighbors.o" "$OUT/obj/corr_kernels.o" "$OUT/obj/corr_tma.o" "$OUT/obj/pgo.o" -lcudart
grep -h -A1 "Compiling entry function\|error\|warning" "$OUT"/obj/*.log | grep -v "^--" | head -80 || true
ls -la "$OUT/libpgba.so"
