#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- not part of the product path.
#
# Compiles the reference's own hot-path kernels, UNMODIFIED and read in place
# from /root/reference (never copied into this repo), into pybind modules under
# oracle/_ref/ (git-ignored, travels to the GPU box with the gpurun snapshot):
#
#   oracle/_ref/_raymarching*.so  <- instance_nerf/raymarching/src/{raymarching.cu,bindings.cpp}
#   oracle/_ref/_gridencoder*.so  <- instance_nerf/gridencoder/src/{gridencoder.cu,bindings.cpp}
#   oracle/_ref/_shencoder*.so    <- instance_nerf/shencoder/src/{shencoder.cu,bindings.cpp}
#
# The reference's own build (backend.py / setup.py) is NOT used: it pins
# -std=c++14 (torch >= 2.1 headers need c++17) and passes no -gencode.
# raymarching.cu additionally needs --expt-relaxed-constexpr (device use of
# std::numeric_limits::max, raymarching.cu:122,134).
#
# Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
# these modules, and only as the checker.
set -euo pipefail
REF=${REF:-/root/reference/instance_nerf}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
mkdir -p "$OUT/obj"
if [ ! -d "$REF" ]; then
  echo "[oracle/build_ref] $REF absent (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
PY=${PYTHON:-python}
TORCH_DIR=$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')
PY_INC=$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')
EXT=$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')
PYBIND_INC=$($PY -c 'import pybind11;print(pybind11.get_include())' 2>/dev/null || echo "$TORCH_DIR/include")
INC="-I$TORCH_DIR/include -I$TORCH_DIR/include/torch/csrc/api/include -I$PY_INC -I$PYBIND_INC -I/usr/local/cuda/include"
ABI=$($PY -c 'import torch;print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')
NVCC_FLAGS="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ --expt-relaxed-constexpr -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=$ABI"
CXX_FLAGS="-O2 -std=c++17 -fPIC -D_GLIBCXX_USE_CXX11_ABI=$ABI"
LIBS="-L$TORCH_DIR/lib -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,$TORCH_DIR/lib"

build_one() {  # $1 = reference op dir, $2 = module name
  local op=$1 mod=$2
  local so="$OUT/$mod$EXT"
  if [ -f "$so" ] && [ "$so" -nt "$REF/$op/src/$op.cu" ]; then echo "[oracle/build_ref] $mod up to date"; return; fi
  nvcc -c "$REF/$op/src/$op.cu" -o "$OUT/obj/$op.o" $NVCC_FLAGS $INC -DTORCH_EXTENSION_NAME=$mod
  g++ -c "$REF/$op/src/bindings.cpp" -o "$OUT/obj/${op}_bindings.o" $CXX_FLAGS $INC -DTORCH_EXTENSION_NAME=$mod
  g++ -shared "$OUT/obj/$op.o" "$OUT/obj/${op}_bindings.o" -o "$so" $LIBS
  echo "[oracle/build_ref] built $so"
}
build_one raymarching _raymarching &
build_one gridencoder _gridencoder &
build_one shencoder  _shencoder &
wait
