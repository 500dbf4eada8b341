#!/usr/bin/env bash
# Build the reference's own CUDA extensions (losses, sampling) for sm_100a from
# the sources where they lie under /root/reference, into oracle/_ref/ (git-ignored,
# shipped to the GPU box by gpurun).  Used as the GPU-side checker and as the
# "reference recompiled for sm_100a" timing arm.  Never copies reference sources.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PP_REFERENCE_ROOT:-/root/reference}/pytorch_points/_ext"
OUT="$HERE/_ref"
[ -d "$REF" ] || { echo "reference sources not found at $REF"; exit 3; }
mkdir -p "$OUT/obj"
PY=python
TI=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'include'))")
TL=$($PY -c "import torch,os;print(os.path.join(os.path.dirname(torch.__file__),'lib'))")
PYI=$($PY -c "import sysconfig;print(sysconfig.get_paths()['include'])")
ABI=$($PY -c "import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))")
COMMON="-O2 -std=c++17 -I$TI -I$TI/torch/csrc/api/include -I$PYI -I$REF -include $HERE/ref_shim.h -D_GLIBCXX_USE_CXX11_ABI=$ABI -DTORCH_API_INCLUDE_EXTENSION_H"
NV="nvcc $COMMON -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr -Xcompiler -fPIC \
 -D__CUDA_NO_HALF_OPERATORS__ -D__CUDA_NO_HALF_CONVERSIONS__ -D__CUDA_NO_BFLOAT16_CONVERSIONS__ -D__CUDA_NO_HALF2_OPERATORS__"
CXX="g++ $COMMON -fPIC -I/usr/local/cuda/include"
LIBS="-L$TL -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$TL"
(
 $NV -DTORCH_EXTENSION_NAME=ref_losses -c "$REF/nmdistance_cuda.cu" -o "$OUT/obj/nmdistance_cuda.o" &
 $NV -DTORCH_EXTENSION_NAME=ref_sampling -c "$REF/sampling_cuda.cu" -o "$OUT/obj/sampling_cuda.o" &
 $NV -DTORCH_EXTENSION_NAME=ref_sampling -c "$REF/interpolate_gpu.cu" -o "$OUT/obj/interpolate_gpu.o" &
 $CXX -DTORCH_EXTENSION_NAME=ref_losses -c "$REF/nmdistance.cpp" -o "$OUT/obj/nmdistance.o" &
 $CXX -DTORCH_EXTENSION_NAME=ref_sampling -c "$REF/sampling.cpp" -o "$OUT/obj/sampling.o" &
 wait
)
g++ -shared -o "$OUT/ref_losses.so" "$OUT/obj/nmdistance.o" "$OUT/obj/nmdistance_cuda.o" $LIBS
g++ -shared -o "$OUT/ref_sampling.so" "$OUT/obj/sampling.o" "$OUT/obj/sampling_cuda.o" "$OUT/obj/interpolate_gpu.o" $LIBS
rm -rf "$OUT/obj"
echo "built $OUT/ref_losses.so $OUT/ref_sampling.so"
