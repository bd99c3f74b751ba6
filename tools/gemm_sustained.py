"""The four projection shapes under the POWER CAP: each kernel runs back to back for --secs
seconds (the first second is discarded), so the rate is what a long prefill phase sustains,
not the burst figure tools/gemm_bench.py prints.  Reports TFLOP/s, median SM clock and power.
cuBLAS (torch.matmul, no epilogue) runs beside each shape for context.
    python tools/gemm_sustained.py [M] [--secs 3]"""
import argparse
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sutro_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("M", nargs="?", type=int, default=32768)
ap.add_argument("--secs", type=float, default=3.0)
ap.add_argument("--shapes", default="qkv,wo,gate_up,down")
ap.add_argument("--variants", default="256,512")
ap.add_argument("--no-cublas", action="store_true")
ap.add_argument("--plain", action="store_true", help="plain store epilogue instead of the shape's own")
a_ = ap.parse_args()
M = a_.M
shapes = [("qkv", 6144, 2560, 0), ("wo", 2560, 4096, 1), ("gate_up", 19456, 2560, 2),
          ("down", 2560, 9728, 1)]

import pynvml  # noqa: E402
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(torch.cuda.current_device())


class Sampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.on, self.mhz, self.w = True, [], []

    def run(self):
        while self.on:
            self.mhz.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.w.append(pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0)
            time.sleep(0.05)


def sustained(fn, flops):
    torch.cuda.synchronize()
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:      # heat-up second, not counted
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
    s = Sampler()
    s.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 0
    e0.record()
    t_end = time.perf_counter() + a_.secs
    while time.perf_counter() < t_end:
        for _ in range(20):
            fn()
        n += 20
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    s.on = False
    ms = e0.elapsed_time(e1) / n
    mhz = sorted(s.mhz)[len(s.mhz) // 2] if s.mhz else 0
    w = sorted(s.w)[len(s.w) // 2] if s.w else 0
    return flops / ms / 1e9, ms * 1e3, mhz, w


tag = " ".join(f"{k[6:]}={v}" for k, v in sorted(os.environ.items()) if k.startswith("SB200_GEMM"))
for name, N, K, epi in shapes:
    if name not in a_.shapes.split(","):
        continue
    if a_.plain:
        epi = 0
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    out = torch.zeros(M, N // 2 if epi == 2 else N, dtype=torch.bfloat16, device="cuda")
    fl = 2.0 * M * N * K
    for bn in [int(v) for v in a_.variants.split(",")]:
        def run():
            L.check(L.lib().sb200_gemm_bf16_tn(L.ptr(a), M, L.ptr(w), L.ptr(out),
                                               L.ptr(out) if epi == 1 else 0, M, N, K,
                                               out.shape[1], epi, bn, L.current_stream()))
        tf, us, mhz, pw = sustained(run, fl)
        print(f"M={M} {name:8s} variant={bn}: {tf:7.1f} TFLOP/s {us:8.1f} us  {mhz} MHz {pw:6.0f} W  {tag}{' plain' if a_.plain else ''}", flush=True)
    if a_.no_cublas:
        continue
    tf, us, mhz, pw = sustained(lambda: torch.matmul(a, w.t()), fl)
    print(f"M={M} {name:8s} cuBLAS     : {tf:7.1f} TFLOP/s {us:8.1f} us  {mhz} MHz {pw:6.0f} W", flush=True)
