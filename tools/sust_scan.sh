#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or qkv_rope" -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
timeout 200 python tools/gemm_sustained.py 32768 --secs 2 --variants 512
