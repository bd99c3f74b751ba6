"""Raster-group scan for the projections with few N tiles (wo, down: N = 2560 = 10 tiles of 256):
one process per group size (SB200_GEMM_GM is read once per process).
    for g in 1 2 3 4 6 9 16 32; do SB200_GEMM_GM=$g python tools/gemm_raster_scan.py; done"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sutro_b200 import _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
shapes = [("qkv", 6144, 2560, 0), ("wo", 2560, 4096, 1), ("down", 2560, 9728, 1)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(8):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


out = []
for name, N, K, epi in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    o = torch.zeros(M, N, dtype=torch.bfloat16, device="cuda")
    for bn in (512, 256):
        ms = timed(lambda: L.check(L.lib().sb200_gemm_bf16_tn(
            L.ptr(a), M, L.ptr(w), L.ptr(o), L.ptr(o) if epi == 1 else 0, M, N, K, N, epi, bn,
            L.current_stream())))
        out.append(f"{name}/{bn}: {2.0 * M * N * K / ms / 1e9:6.0f}")
print(f"GM={os.environ.get('SB200_GEMM_GM', 'auto'):>4s}  " + "  ".join(out), flush=True)
