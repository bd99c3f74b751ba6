"""Diagnostic for the tcgen05 GEMM: structured operands that reveal which (row, k)
elements each output actually consumed.  Writes gpurun_out/gemm_diag.txt."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sutro_b200 import _lib as L  # noqa: E402

os.makedirs("gpurun_out", exist_ok=True)
out_f = open("gpurun_out/gemm_diag.txt", "w")


def log(*a):
    print(*a)
    print(*a, file=out_f, flush=True)


def run(M, N, K, bn, epi=3):
    a = torch.zeros(M, K, dtype=torch.bfloat16, device="cuda")
    # A one-hot: row m selects k = (m*5) % K  -> D[m,n] = W[n, (m*5)%K]
    sel = (torch.arange(M, device="cuda") * 5) % K
    a[torch.arange(M, device="cuda"), sel] = 1
    n_i = torch.arange(N, device="cuda").view(N, 1)
    k_i = torch.arange(K, device="cuda").view(1, K)
    w = ((n_i * 3 + k_i * 7) % 251).to(torch.bfloat16)
    d = torch.full((M, N), -1.0, dtype=torch.float32, device="cuda")
    rc = L.lib().sb200_gemm_bf16_tn(L.ptr(a), M, L.ptr(w), L.ptr(d), 0, M, N, K, N, epi, bn,
                                    L.current_stream())
    if rc:
        log("launch error:", L.lib().sb200_last_error())
        return
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    bad = (d != ref)
    log(f"M={M} N={N} K={K} bn={bn}: mismatches {int(bad.sum())}/{bad.numel()}")
    if bad.any():
        rows = bad.any(1).nonzero().flatten()[:8].tolist()
        cols = bad.any(0).nonzero().flatten()[:8].tolist()
        log("  first bad rows", rows, "first bad cols", cols)
        for r in rows[:4]:
            c = cols[0]
            got = d[r, c:c + 8].tolist()
            want = ref[r, c:c + 8].tolist()
            log(f"  row {r} col {c}: got {got} want {want}")
            # which k would explain got? W[n,k] = (3n+7k)%251
            exp_k = []
            for j, g in enumerate(got):
                ks = [k for k in range(K) if (3 * (c + j) + 7 * k) % 251 == g][:3]
                exp_k.append(ks)
            log(f"    k explaining got: {exp_k}  (selected k={int(sel[r])})")


if __name__ == "__main__":
    torch.cuda.init()
    for (M, N, K, bn) in [(128, 64, 64, 64), (128, 128, 64, 128), (128, 256, 64, 256),
                          (128, 256, 128, 256), (256, 512, 256, 256), (300, 768, 512, 128),
                          (256, 256, 64, 512), (256, 256, 256, 512), (600, 1024, 512, 512)]:
        try:
            run(M, N, K, bn)
        except Exception as e:  # noqa: BLE001
            log("exception", repr(e))
            break
