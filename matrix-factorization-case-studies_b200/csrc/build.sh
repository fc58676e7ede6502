#!/bin/bash
# Builds libcdr_b200.so (sm_100a only) next to the Python package.
set -eo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="${CDR_BUILD_OUT:-$HERE/../convex_dim_red/libcdr_b200.so}"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
SRCS=$(ls "$HERE"/*.cu)
"$NVCC" -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
    -Xcompiler -fPIC -shared ${CDR_NVCC_EXTRA} \
    -o "$OUT" $SRCS
echo "built $OUT"
