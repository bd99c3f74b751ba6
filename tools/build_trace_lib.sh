#!/usr/bin/env bash
# Debug-only twin of libsutro_b200.so whose tcgen05 prefill-attention kernel stamps clock64() at
# its pipeline hand-offs (attn_prefill_tc.cu, SB200_ATTN_TRACE).  Read with tools/attn_trace.py.
set -eu
cd "$(dirname "$0")/../sutro_b200/csrc"
make -s
mkdir -p ../../tools/_trace
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC \
  --expt-relaxed-constexpr -DSB200_ATTN_TRACE -c attn_prefill_tc.cu -o ../../tools/_trace/attn_prefill_tc.o
objs=$(ls build/*.o | grep -v attn_prefill_tc.o)
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../../tools/_trace/libsutro_b200_trace.so \
  $objs ../../tools/_trace/attn_prefill_tc.o -lcudart_static -lpthread -ldl -lrt
echo built tools/_trace/libsutro_b200_trace.so
