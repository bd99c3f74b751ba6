#!/usr/bin/env bash
# sustained-rate scan of experiment knobs on the N = 2560 shapes (one process per setting)
set -u
for c in 74 72 70 64; do for g in 1 2; do
  SB200_GEMM_CLUSTERS=$c SB200_GEMM_GM=$g timeout 120 python tools/gemm_sustained.py 32768 --secs 2 --shapes wo,down --variants 512 --no-cublas
done; done
