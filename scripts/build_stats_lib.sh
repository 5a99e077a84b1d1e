#!/bin/bash
# Diagnostics build: the product library with the express-path / memo counters compiled in
# (-DXLLM_EXP_STATS -DXLLM_MEMO_STATS) -> build/stats/libxllm_ingest_stats.so; use with XLLM_INGEST_LIB=<that path>.
set -e
cd "$(dirname "$0")/.."
make lib >/dev/null
mkdir -p build/stats
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude --expt-relaxed-constexpr \
  -DXLLM_EXP_STATS -c xllm_service_b200/csrc/sp_encode.cu -o build/stats/sp_encode.cu.o
OBJS=$(ls build/*.o | grep -v sp_encode.cu.o)
$NVCC -gencode arch=compute_100a,code=sm_100a -shared -o build/stats/libxllm_ingest_stats.so $OBJS build/stats/sp_encode.cu.o -lcudart -ldl
echo build/stats/libxllm_ingest_stats.so
