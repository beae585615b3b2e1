#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Builds compiled copies of the *reference* into oracle/_ref/ (git-ignored,
# NOT gpurun-ignored, so the .so files travel to the GPU box).  Nothing under /root/reference is
# copied into the repo: sources are compiled where they lie (the stock extension is built from a
# scratch copy under /tmp because setuptools writes into the source tree).
#
#   oracle/_ref/refext/essential_matrix*.so  unmodified reference extension (TORCH_CUDA_ARCH_LIST=10.0a)
#   oracle/_ref/libref_twin_cuda.so          instrumented twin, nvcc, sm_100a (GPU box only)
#   oracle/_ref/libref_host.so               reference solver+cheirality host-compiled with g++
#
# usage: oracle/build_ref.sh [host|twin|ext|all]     (twin and ext take ~10 min each)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF=/root/reference/RANSAC_FiveP
OUT="$HERE/_ref"
WHAT="${1:-all}"
mkdir -p "$OUT"
if [ ! -d "$REF" ]; then echo "reference not present at $REF; nothing to build" >&2; exit 0; fi

build_host() {
  g++ -O2 -std=c++17 -fPIC -shared -w -D__host__= -D__device__= \
      -I"$REF/essential_matrix" -I"$HERE/ref_twin" \
      "$HERE/ref_twin/ref_twin_host.cpp" -o "$OUT/libref_host.so"
  echo "built $OUT/libref_host.so"
}
build_twin() {
  nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -w \
       -gencode arch=compute_100a,code=sm_100a \
       -I"$REF/essential_matrix" -I"$HERE/ref_twin" \
       "$HERE/ref_twin/ref_twin_cuda.cu" -o "$OUT/libref_twin_cuda.so"
  echo "built $OUT/libref_twin_cuda.so"
}
build_ext() {
  local tmp; tmp="$(mktemp -d /tmp/ref_ext_build.XXXXXX)"
  cp -r "$REF" "$tmp/RANSAC_FiveP"
  ( cd "$tmp/RANSAC_FiveP" && TORCH_CUDA_ARCH_LIST=10.0a MAX_JOBS=2 \
      python setup.py -q build_ext --build-lib "$OUT/refext" --build-temp "$tmp/build" )
  rm -rf "$tmp"
  echo "built $(ls "$OUT"/refext/essential_matrix*.so)"
}
case "$WHAT" in
  host) build_host ;;
  twin) build_twin ;;
  ext)  build_ext ;;
  all)  build_host; build_twin & build_ext & wait ;;
esac
