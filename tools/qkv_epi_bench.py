"""Fused QKV GEMM (q/k-norm + RoPE + paged K/V write in the epilogue) at the prefill shape,
with and without the dense K/V copy for the tcgen05 prefill attention (SB200_QKV_DENSE=0/1,
one process each), next to the plain store epilogue.  python tools/qkv_epi_bench.py [M]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sutro_b200 import _lib as L  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
hq, hkv, K = 32, 8, 2560
N = (hq + 2 * hkv) * 128
dev = "cuda"
torch.manual_seed(0)
a = torch.randn(M, K, device=dev).bfloat16()
w = (torch.randn(N, K, device=dev) / math.sqrt(K)).bfloat16()
qn = (torch.rand(128, device=dev) + 0.5).bfloat16()
kn = (torch.rand(128, device=dev) + 0.5).bfloat16()
cos = torch.randn(4096, 64, device=dev).bfloat16()
sin = torch.randn(4096, 64, device=dev).bfloat16()
n_slots = M // 128
tok_slot = (torch.arange(M, dtype=torch.int32, device=dev) // 128)
tok_pos = (torch.arange(M, dtype=torch.int32, device=dev) % 128)
max_pages = 9
pt = torch.arange(n_slots * max_pages, dtype=torch.int32, device=dev).view(n_slots, max_pages)
pool = torch.zeros(n_slots * max_pages, hkv, 2, 16, 128, dtype=torch.bfloat16, device=dev)
out = torch.zeros(M, N, dtype=torch.bfloat16, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(10):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def fused():
    L.check(L.lib().sb200_gemm_qkv_rope(L.ptr(a), M, L.ptr(w), L.ptr(out), M, K, 0, L.ptr(qn),
                                        L.ptr(kn), L.ptr(cos), L.ptr(sin), L.ptr(tok_slot),
                                        L.ptr(tok_pos), L.ptr(pt), max_pages, L.ptr(pool), hq, hkv,
                                        1e-6, L.current_stream()))


def plain():
    L.check(L.lib().sb200_gemm_bf16_tn(L.ptr(a), M, L.ptr(w), L.ptr(out), 0, M, N, K, N, 0, 0,
                                       L.current_stream()))


fl = 2.0 * M * N * K
t = timed(fused)
print(f"SB200_QKV_DENSE={os.environ.get('SB200_QKV_DENSE', '1')} fused QKV M={M}: {t*1e3:.1f} us {fl/t/1e9:.0f} TFLOP/s")
t = timed(plain)
print(f"plain store epilogue            M={M}: {t*1e3:.1f} us {fl/t/1e9:.0f} TFLOP/s")
t = timed(lambda: torch.matmul(a, w.t()))
print(f"cuBLAS (context)                M={M}: {t*1e3:.1f} us {fl/t/1e9:.0f} TFLOP/s")
