"""Test helper: pack / unpack the swizzled paged KV layout (csrc/kernels.h).

pool[page][kv_head][K|V][16 tokens][128 dims] bf16; inside a [16][128] tile the
16-byte chunk c (8 dims) of token row r sits at chunk position c ^ (r & 7).
"""
import torch

PAGE = 16
HD = 128


def _perm(device):
    r = torch.arange(PAGE, device=device).view(PAGE, 1)
    c = torch.arange(16, device=device).view(1, 16)
    return (c ^ (r & 7))  # [16,16]: physical chunk of logical chunk c in row r


def pack_tile(x):
    """x: [..., 16, 128] logical -> physical (swizzled) tile."""
    shp = x.shape
    xc = x.reshape(*shp[:-1], 16, 8)
    out = torch.empty_like(xc)
    p = _perm(x.device)
    idx = p.view(*([1] * (xc.dim() - 3)), PAGE, 16, 1).expand_as(xc)
    out.scatter_(-2, idx, xc)
    return out.reshape(shp)


def unpack_tile(x):
    shp = x.shape
    xc = x.reshape(*shp[:-1], 16, 8)
    p = _perm(x.device)
    idx = p.view(*([1] * (xc.dim() - 3)), PAGE, 16, 1).expand_as(xc)
    return torch.gather(xc, -2, idx).reshape(shp)


def build_cache(k_seqs, v_seqs, hkv, num_pages, device, seed=0):
    """k_seqs[i], v_seqs[i]: [L_i, hkv, 128] bf16.  Returns (pool, page_table, max_pages).
    Pages are assigned in a shuffled order so page ids are non-trivial."""
    g = torch.Generator().manual_seed(seed)
    order = torch.randperm(num_pages, generator=g).tolist()
    n_seq = len(k_seqs)
    max_pages = max((k.shape[0] + PAGE - 1) // PAGE for k in k_seqs) + 1
    pool = torch.zeros(num_pages, hkv, 2, PAGE, HD, dtype=torch.bfloat16, device=device)
    pt = torch.zeros(n_seq, max_pages, dtype=torch.int32)
    nxt = 0
    for i, (k, v) in enumerate(zip(k_seqs, v_seqs)):
        L = k.shape[0]
        for p in range((L + PAGE - 1) // PAGE):
            page = order[nxt]
            nxt += 1
            pt[i, p] = page
            n = min(PAGE, L - p * PAGE)
            kt = torch.zeros(hkv, PAGE, HD, dtype=torch.bfloat16, device=device)
            vt = torch.zeros_like(kt)
            kt[:, :n] = k[p * PAGE:p * PAGE + n].transpose(0, 1)
            vt[:, :n] = v[p * PAGE:p * PAGE + n].transpose(0, 1)
            pool[page, :, 0] = pack_tile(kt)
            pool[page, :, 1] = pack_tile(vt)
    return pool, pt.to(device), max_pages
