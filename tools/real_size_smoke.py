"""Real-size smoke of the other BASELINE.json configurations (parity-test cases, not bench
lines): qwen-3-0.6b generative, qwen-3-embedding-0.6b, llama-3.1-8b with a nested schema.
Checks: outputs validate against the schema; for the 0.6B models, first-decision logits /
embeddings of a few rows agree with the CPU oracle on the same GPU-drawn weights."""
import json
import os
import sys
import time
from typing import List, Literal, Optional

import numpy as np
import torch
from pydantic import BaseModel, Field

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle.bpe_ref import RefTokenizer  # noqa: E402
from oracle.model_ref import RefModel  # noqa: E402
from sutro_b200 import modelspec as MS, synth, vocab as VB  # noqa: E402
from sutro_b200.engine import LocalEngine  # noqa: E402
from sutro_b200.schema_fsm import FsmLimits  # noqa: E402


class Item(BaseModel):
    name: str = Field(max_length=12)
    quantity: int = Field(ge=0, le=1000)
    kind: Literal["a", "b", "c"]
    price: Optional[float] = None


class Order(BaseModel):
    customer: str = Field(max_length=10)
    items: List[Item] = Field(max_length=3)
    paid: bool


def build(name, **kw):
    spec = MS.get_spec(name)
    w = MS.make_engine_weights_on_device(spec, 0, "cuda")
    v = VB.build_vocab(spec.family, spec.vocab_size, seed=0)
    return spec, w, v, LocalEngine(spec, w, v, device=0, kv_pages=4096, **kw)


def main():
    torch.set_num_threads(32)
    # ---- qwen-3-0.6b, generative, G = 2 ----
    spec, w, v, eng = build("qwen-3-0.6b", max_slots=256, max_prefill_tokens=8192)
    rows = synth.product_reviews(512, seed=1)
    t = time.time()
    res = eng.generate(rows, system_prompt=synth.README_SYSTEM_PROMPT, max_new_tokens=16,
                       ignore_eos=True, return_tokens=True, return_first_logits=True)
    print(f"qwen-3-0.6b: {len(rows) / (time.time() - t):.0f} rows/s, stats "
          f"{ {k: res.stats[k] for k in ('prefill_steps', 'decode_steps', 'input_tokens')} }")
    hf = MS.unpack_to_hf(spec, w)
    model, tok = RefModel(spec, hf, fast=True), RefTokenizer(v)
    tpl = VB.chat_template(spec.family, synth.README_SYSTEM_PROMPT)
    for i in range(3):
        want = model.logits(tok.render(tpl, rows[i]))[-1]
        got = res.first_logits[i]
        d = (got - want).abs()
        print(f"  row {i}: logits rms diff {d.pow(2).mean().sqrt() / want.std():.4f} sigma, "
              f"max {d.max() / want.std():.3f} sigma, argmax equal {int(got.argmax()) == int(want.argmax())}")
        assert d.pow(2).mean().sqrt() / want.std() < 0.05
    eng.close()
    del eng, w
    torch.cuda.empty_cache()

    # ---- qwen-3-embedding-0.6b ----
    spec, w, v, eng = build("qwen-3-embedding-0.6b", max_slots=512, max_prefill_tokens=16384)
    rows = synth.short_texts(4096, seed=2)
    t = time.time()
    res = eng.generate(rows)
    print(f"qwen-3-embedding-0.6b: {len(rows) / (time.time() - t):.0f} rows/s, shape {res.embeddings.shape}")
    hf = MS.unpack_to_hf(spec, w)
    model, tok = RefModel(spec, hf, fast=True), RefTokenizer(v)
    tpl = VB.embedding_template(spec.family)
    for i in range(3):
        want = model.embed(tok.render(tpl, rows[i])).numpy()
        print(f"  row {i}: max abs diff {np.abs(res.embeddings[i] - want).max():.4f}, cos "
              f"{float(res.embeddings[i] @ want):.5f}")
        assert float(res.embeddings[i] @ want) > 0.995
    eng.close()
    del eng, w
    torch.cuda.empty_cache()

    # ---- llama-3.1-8b, nested schema ----
    spec, w, v, eng = build("llama-3.1-8b", max_slots=256, max_prefill_tokens=8192)
    rows = synth.extraction_documents(256, seed=3)
    lim = FsmLimits(max_string_chars=12, max_array_items=3)
    t = time.time()
    res = eng.generate(rows, system_prompt="Extract the order as JSON.",
                       json_schema=Order.model_json_schema(), max_new_tokens=160, fsm_limits=lim,
                       return_tokens=True)
    dt = time.time() - t
    ok = cut = 0
    for o, toks in zip(res.outputs, res.out_tokens):
        try:
            Order.model_validate(json.loads(o))
            ok += 1
        except Exception:
            assert len(toks) == 160, (o, len(toks))   # only the token cap may end a row early
            cut += 1
    print(f"llama-3.1-8b nested schema: {len(rows) / dt:.0f} rows/s, {ok}/{len(rows)} outputs validate "
          f"({cut} cut by max_new_tokens), "
          f"fsm_states {res.stats['fsm_states']}, out tokens {res.stats['output_tokens']}")
    print("  sample:", res.outputs[0][:160])
    assert ok + cut == len(rows) and ok > len(rows) // 2
    print("real-size smoke ok")


if __name__ == "__main__":
    main()
