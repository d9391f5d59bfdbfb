#!/usr/bin/env bash
# Builds libpgx.so (the C-ABI library of include/pgx.h) for sm_100a, in-tree.
# -fmad=false: the reference computes d*m + (1-d)*f etc. with separate roundings;
# no FMA contraction keeps max-product results bit-comparable with the oracle.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
nvcc -O3 -std=c++17 -shared -Xcompiler -fPIC -lineinfo -fmad=false \
  -gencode arch=compute_100a,code=sm_100a \
  ${PGX_NVCC_EXTRA:-} \
  -o "$here/libpgx.so" "$here/pgx.cu" -lcudart
echo "built $here/libpgx.so"
