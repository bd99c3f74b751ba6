"""CPU oracle for the transformer forward + greedy decode loop.  TEST INFRASTRUCTURE:
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this; the product path (sutro_b200/) never does.

PARITY UNPINNED by the reference: sutro-sh/sutro contains no local model code (its
infer() POSTs to a hosted service, sutro/sdk.py:195-223), so there
are no reference golden vectors for this path.  The algorithm restated here is the
published one for the named open-weight architectures, as implemented in
transformers 5.5.0 (third-party, pinned by this image):
  * RMSNorm            transformers/models/qwen3/modeling_qwen3.py:50-66
  * SwiGLU MLP         modeling_qwen3.py:70-82
  * rotate-half RoPE   modeling_qwen3.py:151-180 (cos/sin cast to the model dtype first)
  * q/k per-head norm  modeling_qwen3.py:248-264 (before RoPE; Qwen3 only)
  * GQA attention      modeling_qwen3.py:183-210 (fp32 softmax, scale = head_dim**-0.5)
  * decoder layer      modeling_qwen3.py:290-320 (pre-norm, two residual adds)
  * llama3 rope scale  modeling_rope_utils.py:550-625
tests/test_oracle_vs_hf.py pins this file against transformers' own Qwen3ForCausalLM /
LlamaForCausalLM run on the same seeded weights (both importable in this image and on
the GPU box), which is the strongest pin available.

Numerics: values are stored in bf16 and every matmul accumulates in fp32, with a
rounding to bf16 at each op boundary where a bf16 PyTorch model rounds (linear
outputs, norm outputs, each RoPE multiply/add, SiLU, gate*up, residual adds).
Attention probabilities are rounded to bf16 before P·V; logits stay fp32 so the
arg-max is not decided by bf16 ties.  These are the engine's rounding points too.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

BF = torch.bfloat16


_EXACT = False   # RefModel.logits(..., exact=True): no intermediate rounding (see there)


def _r(x: torch.Tensor) -> torch.Tensor:
    """round to bf16, keep computing in fp32"""
    return x if _EXACT else x.to(BF).float()


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    xf = x.float()
    var = xf.pow(2).mean(-1, keepdim=True)
    return _r(w.float() * _r(xf * torch.rsqrt(var + eps)))


def linear(x: torch.Tensor, w: torch.Tensor, fast: bool = False) -> torch.Tensor:
    """bf16 x bf16 -> fp32 accumulate -> bf16 (returned as fp32 holding bf16 values).
    `fast`: the weight was pre-converted to fp32 once (RefModel(fast=True)), so the product
    is a plain sgemm — the same contract without re-converting 8 GB of weights per call."""
    if fast:
        return _r(x @ w.t())
    return _r(x.float() @ w.float().t())


def rope(x: torch.Tensor, cos: torch.Tensor, sin: torch.Tensor) -> torch.Tensor:
    """x: [T,H,128] fp32 holding bf16 values; cos/sin: [T,64] bf16.  bf16 op by op."""
    c = torch.cat([cos, cos], -1).float()[:, None, :]
    s = torch.cat([sin, sin], -1).float()[:, None, :]
    half = x.shape[-1] // 2
    rot = torch.cat([-x[..., half:], x[..., :half]], -1)
    return _r(_r(x * c) + _r(rot * s))


@dataclass
class GenResult:
    tokens: List[int]
    margins: List[float]        # top1 - top2 logit among allowed tokens, one per DECISION
                                # (forced jump-forward tokens are not decisions);
                                # decision_index[i] = position in `tokens` of decision i
    finished_by: str            # "eos" | "fsm" | "length"
    runner_up: Optional[List[int]] = None   # second-best allowed token at each decision
    decision_pos: Optional[List[int]] = None  # index into `tokens` of each decision (EOS: len)
    near: Optional[List[Dict[int, float]]] = None  # per decision: {token: top1 - logit} of the
                                                   # (up to 8) best allowed tokens


class RefModel:
    def __init__(self, spec, weights: Dict[str, torch.Tensor], fast: bool = False):
        from sutro_b200.modelspec import rope_tables
        self.spec = spec
        self.fast = fast
        if fast:  # big-model CPU baseline: keep every matrix in fp32 (bf16 values) up front
            weights = {k: (v.float() if v.dim() == 2 and "embed_tokens" not in k else v)
                       for k, v in weights.items()}
        self.w = weights
        self.cos, self.sin = rope_tables(spec)
        self._lm = None

    # -- pieces -----------------------------------------------------------
    def _lm_head(self) -> torch.Tensor:
        if self._lm is None:
            key = "model.embed_tokens.weight" if self.spec.tied_embeddings else "lm_head.weight"
            self._lm = self.w[key].float()
        return self._lm

    def _layer(self, i: int, x: torch.Tensor, pos: torch.Tensor, cache) -> torch.Tensor:
        sp, w, p = self.spec, self.w, f"model.layers.{i}."
        T = x.shape[0]
        h = rmsnorm(x, w[p + "input_layernorm.weight"], sp.rms_eps)
        q = linear(h, w[p + "self_attn.q_proj.weight"], self.fast).view(T, sp.n_q_heads, -1)
        k = linear(h, w[p + "self_attn.k_proj.weight"], self.fast).view(T, sp.n_kv_heads, -1)
        v = linear(h, w[p + "self_attn.v_proj.weight"], self.fast).view(T, sp.n_kv_heads, -1)
        if sp.qk_norm:
            q = rmsnorm(q, w[p + "self_attn.q_norm.weight"], sp.rms_eps)
            k = rmsnorm(k, w[p + "self_attn.k_norm.weight"], sp.rms_eps)
        q = rope(q, self.cos[pos], self.sin[pos])
        k = rope(k, self.cos[pos], self.sin[pos])
        if cache[i] is None:
            cache[i] = (k, v)
        else:
            cache[i] = (torch.cat([cache[i][0], k], 0), torch.cat([cache[i][1], v], 0))
        K, V = cache[i]                                   # [L, hkv, 128]
        g = sp.n_q_heads // sp.n_kv_heads
        Kh = K.repeat_interleave(g, dim=1)                # [L, hq, 128]
        Vh = V.repeat_interleave(g, dim=1)
        s = torch.einsum("thd,lhd->htl", q, Kh) * (sp.head_dim ** -0.5)
        L = K.shape[0]
        kpos = torch.arange(L)[None, None, :]
        s = s.masked_fill(kpos > pos[None, :, None], float("-inf"))
        pr = _r(torch.softmax(s, dim=-1))
        a = _r(torch.einsum("htl,lhd->thd", pr, Vh)).reshape(T, -1)
        x = _r(linear(a, w[p + "self_attn.o_proj.weight"], self.fast) + x)
        h = rmsnorm(x, w[p + "post_attention_layernorm.weight"], sp.rms_eps)
        gate = linear(h, w[p + "mlp.gate_proj.weight"], self.fast)
        up = linear(h, w[p + "mlp.up_proj.weight"], self.fast)
        act = _r(_r(F.silu(gate)) * up)
        return _r(linear(act, w[p + "mlp.down_proj.weight"], self.fast) + x)

    def _forward(self, ids: Sequence[int], start: int, cache) -> torch.Tensor:
        """Run tokens `ids` at positions start.. through the stack; returns the final
        normed hidden states [T, d] (fp32 holding bf16 values)."""
        pos = torch.arange(start, start + len(ids))
        x = self.w["model.embed_tokens.weight"][torch.tensor(list(ids))].float()
        for i in range(self.spec.n_layers):
            x = self._layer(i, x, pos, cache)
        return rmsnorm(x, self.w["model.norm.weight"], self.spec.rms_eps)

    # -- public -----------------------------------------------------------
    @torch.no_grad()
    def logits(self, ids: Sequence[int], exact: bool = False) -> torch.Tensor:
        """Teacher-forced fp32 logits for every position: [T, V].  `exact=True` computes the
        same network on the same bf16-valued weights WITHOUT the intermediate bf16 roundings
        (fp32 throughout): the yardstick for how much of an engine-vs-oracle difference is the
        rounding noise every bf16 pipeline carries (tests/test_zy_real_size_parity_gpu.py)."""
        global _EXACT
        old, _EXACT = _EXACT, bool(exact)
        try:
            h = self._forward(ids, 0, [None] * self.spec.n_layers)
            return h @ self._lm_head().t()
        finally:
            _EXACT = old

    @torch.no_grad()
    def embed(self, ids: Sequence[int]) -> torch.Tensor:
        """Embedding-model head: last-token hidden state, L2 normalised, fp32 [d]."""
        h = self._forward(ids, 0, [None] * self.spec.n_layers)[-1]
        return h / h.norm().clamp_min(1e-12)

    @torch.no_grad()
    def generate(self, prompt: Sequence[int], max_new: int, eos_id: int,
                 ignore_eos: bool = False,
                 fsm=None, tok_bytes: Optional[Callable[[int], bytes]] = None) -> GenResult:
        """Greedy decode.  `fsm` (oracle/fsm_ref.TokenFSM-like) supplies, per state, a bool
        mask over the vocabulary and the transition on a token."""
        cache = [None] * self.spec.n_layers
        forced = list(getattr(fsm, "forced_prefix", [])) if fsm is not None else []
        prompt = list(prompt) + forced          # jump-forward: forced output rides with the prompt
        h = self._forward(prompt, 0, cache)[-1:]
        out: List[int] = list(forced)
        margins: List[float] = []
        second: List[int] = []
        dpos: List[int] = []
        near: List[Dict[int, float]] = []
        state = fsm.start if fsm is not None else None
        pos = len(prompt)
        why = "length"
        while True:
            lg = (h @ self._lm_head().t())[0]
            if fsm is not None:
                allowed = fsm.allowed(state)
                lg = torch.where(allowed, lg, torch.full_like(lg, float("-inf")))
            top = torch.topk(lg, min(8, lg.numel()))
            near.append({int(i): float(top.values[0] - v) for v, i in zip(top.values, top.indices)
                         if v > float("-inf")})
            # lowest index wins exact ties, like the engine's sampler
            best = top.values[0]
            tok = int((lg == best).nonzero()[0])
            margins.append(float(top.values[0] - top.values[1]))
            dpos.append(len(out))
            second.append(int(top.indices[1]) if int(top.indices[0]) == tok else int(top.indices[0]))
            if tok == eos_id and not ignore_eos:
                why = "eos"
                break
            out.append(tok)
            if fsm is not None:
                state = fsm.step(state, tok)
                if fsm.is_final(state):
                    why = "fsm"
                    break
                tail = getattr(fsm, "tails", {}).get(state)
                if tail:                         # forced all the way to a final state
                    out += tail[:max(0, max_new - len(out))]
                    why = "fsm"
                    break
            if len(out) >= max_new:
                break
            h = self._forward([tok], pos, cache)
            pos += 1
        return GenResult(out, margins, why, second, dpos, near)
