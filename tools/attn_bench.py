#!/usr/bin/env python
"""Prefill attention in isolation: the tcgen05 dense kernel (attn_prefill_tc.cu) against the
paged mma.sync kernel (attn_prefill.cu) on the benchmark's shape — one prefill step of
~32 k new tokens, sequences of `--new` tokens behind a `--past`-token shared prefix,
qwen-3-4b heads (32 q / 8 kv).  Prints ms per launch and the effective rates.

    python tools/attn_bench.py [--new 110] [--past 46] [--tokens 32768]
"""
import argparse
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from sutro_b200 import _lib as L  # noqa: E402
import kv_layout as KV  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--new", type=int, default=110)
ap.add_argument("--past", type=int, default=46)
ap.add_argument("--tokens", type=int, default=32768)
ap.add_argument("--hq", type=int, default=32)
ap.add_argument("--hkv", type=int, default=8)
ap.add_argument("--iters", type=int, default=20)
ap.add_argument("--trace", type=int, default=-1, help="CTA whose pipeline timeline to print "
                "(needs SB200_LIB=tools/_trace/libsutro_b200_trace.so)")
ap.add_argument("--trace-from", type=int, default=40)
ap.add_argument("--trace-n", type=int, default=12)
a = ap.parse_args()
dev = "cuda"
hq, hkv = a.hq, a.hkv
n_seq = a.tokens // a.new
T = n_seq * a.new
ldq = (hq + 2 * hkv) * 128
torch.manual_seed(0)
qkv = (torch.randn(T + 64, ldq, device=dev) * 0.5).to(torch.bfloat16)
pre = (torch.randn(1, max(a.past, 1), 2 * hkv * 128, device=dev) * 0.5).to(torch.bfloat16)
out = torch.zeros(T, hq * 128, dtype=torch.bfloat16, device=dev)
out2 = torch.zeros_like(out)
qt = L.lib().sb200_attn_prefill_q_tile(hq, hkv)
items, q_start = [], []
for i in range(n_seq):
    q_start.append(i * a.new)
    for t0 in range(0, a.new, qt):
        items += [i, t0]
i32 = lambda x: torch.tensor(x, dtype=torch.int32, device=dev)
d_items, d_qs = i32(items), i32(q_start)
d_ql, d_past = i32([a.new] * n_seq), i32([a.past] * n_seq)
scale = 1.0 / math.sqrt(128)
st = L.current_stream()

# paged copy of the same K/V for the mma.sync kernel: prefix pages are shared
P = KV.PAGE
n_pre_pages = a.past // P
own_pages = (a.past + a.new + P - 1) // P - n_pre_pages
num_pages = n_pre_pages + n_seq * own_pages + 1
max_pages = n_pre_pages + own_pages + 1
pool = torch.zeros(num_pages, hkv, 2, P, 128, dtype=torch.bfloat16, device=dev)
pt = torch.zeros(n_seq, max_pages, dtype=torch.int32)
rows = qkv[:T].view(n_seq, a.new, hq + 2 * hkv, 128)
pk = pre[0, :a.past, :hkv * 128].view(a.past, hkv, 128)
pv = pre[0, :a.past, hkv * 128:].view(a.past, hkv, 128)
for pg in range(n_pre_pages):
    pool[pg, :, 0] = KV.pack_tile(pk[pg * P:(pg + 1) * P].transpose(0, 1).contiguous())
    pool[pg, :, 1] = KV.pack_tile(pv[pg * P:(pg + 1) * P].transpose(0, 1).contiguous())
kall = torch.cat([pk[None].expand(n_seq, -1, -1, -1), rows[:, :, hq:hq + hkv]], 1)   # [n, past+new, hkv, 128]
vall = torch.cat([pv[None].expand(n_seq, -1, -1, -1), rows[:, :, hq + hkv:]], 1)
tail0 = n_pre_pages * P
L_tail = a.past + a.new - tail0
kt = torch.zeros(n_seq, own_pages * P, hkv, 128, dtype=torch.bfloat16, device=dev)
vt = torch.zeros_like(kt)
kt[:, :L_tail] = kall[:, tail0:]
vt[:, :L_tail] = vall[:, tail0:]
kt = kt.view(n_seq, own_pages, P, hkv, 128).permute(0, 1, 3, 2, 4).contiguous()
vt = vt.view(n_seq, own_pages, P, hkv, 128).permute(0, 1, 3, 2, 4).contiguous()
base = n_pre_pages
pool[base:base + n_seq * own_pages, :, 0] = KV.pack_tile(kt.view(-1, hkv, P, 128))
pool[base:base + n_seq * own_pages, :, 1] = KV.pack_tile(vt.view(-1, hkv, P, 128))
for i in range(n_seq):
    pt[i, :n_pre_pages] = torch.arange(n_pre_pages)
    pt[i, n_pre_pages:n_pre_pages + own_pages] = base + i * own_pages + torch.arange(own_pages)
d_pt, d_slot = pt.to(dev), i32(list(range(n_seq)))


def run_tc():
    L.check(L.lib().sb200_attn_prefill_dense(
        L.ptr(qkv), qkv.shape[0], L.ptr(out), L.ptr(pre) if a.past else None, a.past, 1, 0,
        L.ptr(d_items), len(items) // 2, L.ptr(d_qs), L.ptr(d_ql), L.ptr(d_past), hq, hkv, scale, st))


def run_v1():
    L.check(L.lib().sb200_attn_prefill(
        L.ptr(qkv), L.ptr(out2), L.ptr(pool), L.ptr(d_pt), max_pages, L.ptr(d_items),
        len(items) // 2, L.ptr(d_slot), L.ptr(d_qs), L.ptr(d_ql), L.ptr(d_past), hq, hkv, scale, st))


def timeit(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / a.iters


if a.trace >= 0:
    import ctypes as C
    CAP = 1024
    buf = torch.zeros(6 * CAP, dtype=torch.int64, device=dev)
    fn = L.lib().sb200_attn_trace
    fn.argtypes, fn.restype = [C.c_void_p, C.c_int], C.c_int
    for _ in range(3):
        run_tc()
    torch.cuda.synchronize()
    assert fn(L.ptr(buf), a.trace) == 0
    run_tc()
    torch.cuda.synchronize()
    fn(None, 0)
    tr = buf.cpu().view(6, CAP).tolist()   # 32-bit %clock stamps: fine within one launch
    g0 = a.trace_from
    t0 = tr[2][4 * g0]
    rel = lambda x: "     ." if x == 0 else f"{x - t0:6d}"
    print("cycles relative to the MMA warp reaching block", g0, "(CTA", a.trace, ")")
    print("  g | K: slot  issued | V: slot  issued | S-MMA: at   kfull  issue | PV: at   pfull  vfull  issue |"
          " softmax: at  sfull  ld'd   math  pfree  p-out  epi | max'd  alpha")
    for g in range(g0, g0 + a.trace_n):
        k, v, m, pv, sm = (tr[0][2 * g:2 * g + 2], tr[1][2 * g:2 * g + 2], tr[2][4 * g:4 * g + 3],
                           tr[3][4 * g:4 * g + 4], tr[4][8 * g:8 * g + 7])
        new = "*" if tr[2][4 * g + 3] else " "
        print(f"{g:3d}{new}| {rel(k[0])} {rel(k[1])} | {rel(v[0])} {rel(v[1])} | {rel(m[0])} {rel(m[1])} {rel(m[2])} |"
              f" {rel(pv[0])} {rel(pv[1])} {rel(pv[2])} {rel(pv[3])} | " + " ".join(rel(x) for x in sm) +
              " | " + " ".join(rel(x) for x in tr[5][4 * g:4 * g + 2]))
    last = max(i for i in range(CAP // 8) if tr[4][8 * i + 5])
    print("blocks:", last + 1, " cycles per block:", (tr[4][8 * last + 5] - tr[4][5]) // max(last, 1))
    sys.exit(0)

t_tc, t_v1 = timeit(run_tc), timeit(run_v1)
diff = (out.float() - out2.float()).abs().max().item()
ctx = a.past + (a.new + 1) / 2
flops = 4.0 * T * ctx * 128 * hq          # QK^T + PV, causal
bytes_ = T * (hq * 128 * 2 * 2 + 2 * hkv * 128 * 2)   # q + o + own k,v once
print(f"shape: {n_seq} seqs x {a.new} new tokens behind {a.past} prefix, T={T}, items={len(items)//2*hkv}")
print(f"tcgen05 dense : {t_tc:.3f} ms  {flops/t_tc/1e9:.0f} TFLOP/s(causal)  {bytes_/t_tc/1e6:.0f} GB/s")
print(f"mma.sync paged: {t_v1:.3f} ms  {flops/t_v1/1e9:.0f} TFLOP/s(causal)  {bytes_/t_v1/1e6:.0f} GB/s")
print(f"max |tc - v1| = {diff:.4f}")
