#!/bin/bash
# Build an A/B variant of libcidb200.so with extra nvcc defines:  tools/build_variant.sh <name> -DCID_ATTN_V3 ...
# -> tools/bin/libcidb200_<name>.so ; select it with CID_LIB_PATH=tools/bin/libcidb200_<name>.so (same ABI, consistentid_b200/lib.py).
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p tools/bin
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 --shared -Xcompiler -fPIC "$@" \
  -o tools/bin/libcidb200_${name}.so consistentid_b200/csrc/api.cu
echo built tools/bin/libcidb200_${name}.so
