#!/usr/bin/env bash
# sustained-rate scan of the raster group size on the N = 2560 shapes (one process per setting)
set -u
for g in 1 2 4 8 16 32; do
  SB200_GEMM_GM=$g timeout 120 python tools/gemm_sustained.py 32768 --secs 2 --shapes wo,down --variants 512 --no-cublas
done
