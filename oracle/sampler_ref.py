"""CPU oracle for temperature / top-k / top-p sampling.  TEST INFRASTRUCTURE (see
oracle/model_ref.py for who may import oracle/).

PARITY UNPINNED by the reference (`sampling_params` is an opaque dict forwarded to the
hosted service, sutro/sdk.py:203).  Semantics restated here are the customary ones
(vLLM / transformers logits processors): z = logit / T over the allowed tokens; top-k keeps
every token whose z is >= the k-th largest (ties kept); top-p then keeps the smallest set of
highest-z tokens whose softmax mass reaches top_p (ties kept); the token is drawn by inverse
CDF over the kept set in vocabulary order from one uniform u per (seed, row, step).
The generator is Philox4x32-10 (Salmon et al., SC'11), pinned by the Random123 known-answer
vectors in tests/test_sampler_oracle.py; the kept sets are checked there against
transformers' Temperature / TopK / TopP logits warpers.
"""
from __future__ import annotations

import numpy as np

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(key, ctr):
    k0, k1 = key
    c0, c1, c2, c3 = ctr
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> 32, p0 & MASK, p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def uniform(seed: int, row: int, step: int, seed_per_row: bool) -> float:
    r = row if seed_per_row else 0
    x = philox4x32_10((seed & MASK, (seed >> 32) & MASK), (r & MASK, (r >> 32) & MASK, step, 0))
    return float(np.float32(x[0] >> 8) * np.float32(1.0 / 16777216.0))


def kept_set(logits: np.ndarray, allowed: np.ndarray, temperature: float, top_k: int,
             top_p: float):
    """-> (kept bool mask, weights exp(z - zmax) as float32)"""
    z = (logits.astype(np.float32) * np.float32(1.0 / temperature)).astype(np.float32)
    z = np.where(allowed, z, -np.inf).astype(np.float32)
    zmax = z.max()
    w = np.exp(z - zmax, dtype=np.float32)
    w[~allowed] = 0
    keep = allowed.copy()
    if top_k and top_k > 0 and allowed.sum() > top_k:
        kth = np.sort(z[allowed])[-top_k]
        keep &= z >= kth
    if 0.0 < top_p < 1.0:
        zk = np.where(keep, z, -np.inf)
        order = np.argsort(-zk, kind="stable")
        mass = np.cumsum(w[order].astype(np.float64))
        total = mass[keep.sum() - 1]
        n = int(np.searchsorted(mass[:keep.sum()], top_p * total, side="left")) + 1
        thr = zk[order[min(n, keep.sum()) - 1]]
        keep &= z >= thr
    return keep, w


def sample(logits: np.ndarray, allowed: np.ndarray, temperature: float, top_k: int, top_p: float,
           u: float):
    """-> (token, slack): slack = distance of the draw from the nearest CDF boundary in units
    of total mass (a GPU/CPU disagreement is only acceptable when slack is ~fp32 rounding)."""
    keep, w = kept_set(logits, allowed, temperature, top_k, top_p)
    wk = np.where(keep, w, 0).astype(np.float64)
    cdf = np.cumsum(wk)
    total = cdf[-1]
    target = u * total
    tok = int(np.searchsorted(cdf, target, side="right"))
    tok = min(tok, len(cdf) - 1)
    while wk[tok] == 0 and tok > 0:          # target == total lands past the last kept token
        tok -= 1
    lo = cdf[tok] - wk[tok]
    slack = min(target - lo, cdf[tok] - target) / total
    return tok, float(slack)


def logprob(logits: np.ndarray, allowed: np.ndarray, temperature: float, tok: int) -> float:
    t = temperature if temperature > 0 else 1.0
    z = np.where(allowed, logits.astype(np.float64) / t, -np.inf)
    zmax = z.max()
    return float(z[tok] - zmax - np.log(np.exp(z - zmax).sum()))
