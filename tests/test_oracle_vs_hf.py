"""Pins oracle/model_ref.py against transformers' own model classes on the same seeded
weights (the reference repo holds no model code or golden vectors — SURVEY.md §8c)."""
import pytest
import torch

from oracle.model_ref import RefModel
from sutro_b200 import modelspec as MS


def hf_model(spec, weights):
    import transformers
    common = dict(vocab_size=spec.vocab_size, hidden_size=spec.d_model,
                  intermediate_size=spec.d_ff, num_hidden_layers=spec.n_layers,
                  num_attention_heads=spec.n_q_heads, num_key_value_heads=spec.n_kv_heads,
                  head_dim=spec.head_dim, max_position_embeddings=spec.max_position,
                  rms_norm_eps=spec.rms_eps, tie_word_embeddings=spec.tied_embeddings,
                  attention_bias=False, attn_implementation="eager")
    if spec.family == "qwen3":
        cfg = transformers.Qwen3Config(rope_theta=spec.rope_theta, **common)
        cls = transformers.Qwen3ForCausalLM
    else:
        rp = {"rope_type": "llama3", "rope_theta": spec.rope_theta, **spec.rope_scaling}
        cfg = transformers.LlamaConfig(rope_parameters=rp, mlp_bias=False, **common)
        cls = transformers.LlamaForCausalLM
    m = cls(cfg).eval()
    sd = {k: v.float() for k, v in weights.items()}
    if spec.tied_embeddings:
        sd["lm_head.weight"] = sd["model.embed_tokens.weight"]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all("rotary" in k or "inv_freq" in k for k in missing), missing
    return m


@pytest.mark.parametrize("name", ["tiny-qwen3", "tiny-qwen3-g4", "tiny-llama"])
def test_oracle_logits_match_transformers(name):
    spec = MS.get_spec(name)
    w = MS.make_weights(spec, seed=3, std=0.05)
    ids = torch.randint(0, spec.vocab_size, (40,), generator=torch.Generator().manual_seed(1))
    ref = RefModel(spec, w).logits(ids.tolist())
    m = hf_model(spec, w)
    with torch.no_grad():
        hf32 = m(ids[None]).logits[0]
    scale = hf32.std().item()
    # fp32 transformers vs bf16-rounded oracle: bf16 noise only
    assert (ref - hf32).abs().max().item() < 0.1 * scale
    assert (ref - hf32).abs().mean().item() < 0.015 * scale
    # bf16 transformers rounds at the same op boundaries as the oracle
    with torch.no_grad():
        hf16 = m.to(torch.bfloat16)(ids[None]).logits[0].float()
    assert (ref - hf16).abs().max().item() < 0.1 * scale
    # decisions agree wherever the fp32 model's top-1 margin is not tiny
    top2 = hf32.topk(2, dim=-1).values
    decided = (top2[:, 0] - top2[:, 1]) > 0.05 * scale
    assert decided.float().mean() > 0.5
    assert torch.equal(ref.argmax(-1)[decided], hf32.argmax(-1)[decided])


def test_rope_inv_freq_matches_transformers_llama3():
    spec = MS.get_spec("tiny-llama")
    m = hf_model(spec, MS.make_weights(spec, seed=0))
    assert torch.equal(m.model.rotary_emb.inv_freq.float(), MS.rope_inv_freq(spec))


def test_fast_mode_is_the_same_model():
    """RefModel(fast=True) (fp32-cached weights, used by the CPU baseline) == default."""
    spec = MS.get_spec("tiny-qwen3")
    w = MS.make_weights(spec, seed=1, std=0.05)
    ids = list(range(5, 40))
    a = RefModel(spec, w).logits(ids)
    b = RefModel(spec, w, fast=True).logits(ids)
    assert torch.equal(a, b)
