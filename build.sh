#!/usr/bin/env bash
# Builds the C-ABI shared library in-tree: sopro_b200/lib/libsopro_b200.so (sm_100a only).
set -euo pipefail
cd "$(dirname "$0")"
mkdir -p sopro_b200/lib
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared
       -Xptxas -v --expt-relaxed-constexpr)
SRCS=(sopro_b200/csrc/ar_engine.cu)
[ -f sopro_b200/csrc/mimi_engine.cu ] && SRCS+=(sopro_b200/csrc/mimi_engine.cu)
[ -f sopro_b200/csrc/nar_engine.cu ] && SRCS+=(sopro_b200/csrc/nar_engine.cu)
[ -f sopro_b200/csrc/noise_host.cu ] && SRCS+=(sopro_b200/csrc/noise_host.cu)
"$NVCC" "${FLAGS[@]}" -o sopro_b200/lib/libsopro_b200.so "${SRCS[@]}" 2>&1 | tee sopro_b200/lib/build.log | grep -E "error|warning|spill|registers" | sort | uniq -c | sort -rn | head -40
echo "built sopro_b200/lib/libsopro_b200.so"
