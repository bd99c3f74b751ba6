"""Stand-alone timing of the tcgen05 GEMM variants on the engine's shapes (CUDA events,
L2 flushed between iterations).  Usage: python tools/gemm_bench.py [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sutro_b200 import _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
shapes = [("qkv", 6144, 2560, 0), ("wo", 2560, 4096, 1), ("gate_up", 19456, 2560, 2),
          ("down", 2560, 9728, 1)]
variants = [int(v) for v in os.environ.get("VARIANTS", "256,512,516,514").split(",")]
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


for name, N, K, epi in shapes:
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    out = torch.zeros(M, N // 2 if epi == 2 else N, dtype=torch.bfloat16, device="cuda")
    for bn in variants:
        def run():
            L.check(L.lib().sb200_gemm_bf16_tn(L.ptr(a), M, L.ptr(w), L.ptr(out),
                                               L.ptr(out) if epi == 1 else 0, M, N, K,
                                               out.shape[1], epi, bn, L.current_stream()))
        ms = timed(run)
        print(f"M={M} {name:8s} N={N:6d} K={K:5d} variant={bn}: {ms * 1e3:8.1f} us  "
              f"{2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
    ms = timed(lambda: torch.matmul(a, w.t()))
    print(f"M={M} {name:8s} cuBLAS via torch.matmul (context only): {ms * 1e3:8.1f} us  "
          f"{2.0 * M * N * K / ms / 1e9:7.1f} TFLOP/s", flush=True)
