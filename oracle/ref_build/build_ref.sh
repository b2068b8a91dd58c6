#!/usr/bin/env bash
# Compile the reference's own CUDA kernels, source-unmodified and by path from /root/reference, into
# oracle/_ref/{ref_cuda_ba,ref_cuda_corr}*.so for use as a live GPU-side oracle and "reference kernels on the
# same B200" timing.  Outputs only under oracle/_ref/ (git-ignored, shipped
# by gpurun).  No reference source is copied into the repo.  Test infrastructure only.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../_ref"
REF="${REFERENCE_ROOT:-/root/reference}"
[ -d "$REF/cdvslam/fastba" ] || { echo "reference tree not present; keeping prebuilt oracle/_ref"; exit 0; }
mkdir -p "$OUT/obj"
PY=python
TORCH_INC=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
TORCH_LIB=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
PY_INC=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
EXT=$($PY -c "import sysconfig;print(sysconfig.get_config_var('EXT_SUFFIX'))")
INC="-I$TORCH_INC -I$TORCH_INC/torch/csrc/api/include -I/usr/local/cuda/include -I$PY_INC"
NVF="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC -w"
CXF="-O2 -std=c++17 -fPIC -w -D_GLIBCXX_USE_CXX11_ABI=1"
LIBS="-L$TORCH_LIB -ltorch -ltorch_cpu -ltorch_cuda -lc10 -lc10_cuda -ltorch_python -L/usr/local/cuda/lib64 -lcudart"

build_ba() {
  nvcc -c $NVF $INC -DTORCH_EXTENSION_NAME=ref_cuda_ba "$REF/cdvslam/fastba/ba_cuda.cu" -o "$OUT/obj/ba_cuda.o" &
  nvcc -c $NVF $INC -I"$HERE/shim" -DTORCH_EXTENSION_NAME=ref_cuda_ba "$REF/cdvslam/fastba/block_e.cu" -o "$OUT/obj/block_e.o" &
  g++ -c $CXF $INC -DTORCH_EXTENSION_NAME=ref_cuda_ba "$HERE/ref_binding_ba.cpp" -o "$OUT/obj/bind_ba.o" &
  wait
  g++ -shared "$OUT/obj/ba_cuda.o" "$OUT/obj/block_e.o" "$OUT/obj/bind_ba.o" $LIBS -Wl,-rpath,"$TORCH_LIB" -o "$OUT/ref_cuda_ba$EXT"
}
build_corr() {
  nvcc -c $NVF $INC -include "$HERE/shim/type_dispatch_shim.h" -DTORCH_EXTENSION_NAME=ref_cuda_corr \
      "$REF/cdvslam/altcorr/correlation_kernel.cu" -o "$OUT/obj/correlation_kernel.o" &
  g++ -c $CXF $INC -DTORCH_EXTENSION_NAME=ref_cuda_corr "$HERE/ref_binding_corr.cpp" -o "$OUT/obj/bind_corr.o" &
  wait
  g++ -shared "$OUT/obj/correlation_kernel.o" "$OUT/obj/bind_corr.o" $LIBS -Wl,-rpath,"$TORCH_LIB" -o "$OUT/ref_cuda_corr$EXT"
}
build_ba
build_corr
ls -la "$OUT"
