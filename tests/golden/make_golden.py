"""Generates the committed golden fixtures from the third-party libraries that define the
algorithms the oracle restates (the reference repo has no local implementation of this path
and therefore no golden vectors of its own — SURVEY.md §8c).

    python tests/golden/make_golden.py        # rewrites tests/golden/*.npz / *.json

  hf_logits_<model>.npz   transformers 5.5.0 fp32 logits (last 4 positions) of the seeded tiny
                          architectures on a fixed id sequence
  tokenizer_ids.json      tokenizers 0.22.2 ids of the fixture texts under the seeded
                          synthetic vocabularies
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from sutro_b200 import modelspec as MS, synth, vocab as V  # noqa: E402

TEXTS = synth.product_reviews(12, seed=5) + [
    "", " ", "don't DON'T we'LL they've I'm he'd it's", "tabs\tand\nnewlines\r\n\r\n  x  \n\n",
    "numbers 1234567 3.14159 v2.1", "punct!!! ... ((nested)) [x]{y} #tag @user",
    "unicode: café naïve Straße 東京 Привет мир ١٢٣ ½ 🙂🙂 end", "aaaaaaaaaaaaaaaa abababab",
    "mixed123abc456 7z", "\n\n a", "''s 's it 's O'Re", "trailing   "]


def main():
    import transformers
    import tokenizers
    from test_oracle_vs_hf import hf_model
    for name in ("tiny-qwen3", "tiny-qwen3-g4", "tiny-llama"):
        spec = MS.get_spec(name)
        w = MS.make_weights(spec, seed=3, std=0.05)
        ids = torch.randint(0, spec.vocab_size, (40,), generator=torch.Generator().manual_seed(1))
        with torch.no_grad():
            lg = hf_model(spec, w)(ids[None]).logits[0, -4:].float().numpy()
        np.savez_compressed(os.path.join(HERE, f"hf_logits_{name}.npz"), ids=ids.numpy(),
                            logits_last4=lg.astype(np.float32),
                            transformers=np.array(transformers.__version__))
    out = {"tokenizers": tokenizers.__version__, "texts": TEXTS, "ids": {}}
    for fam, size in (("qwen3", 2048), ("llama", 1024)):
        v = V.build_vocab(fam, size, seed=0, n_trained=600)
        hf = V.to_hf_tokenizer(v)
        out["ids"][f"{fam}-{size}"] = [hf.encode(t, add_special_tokens=False).ids for t in TEXTS]
    json.dump(out, open(os.path.join(HERE, "tokenizer_ids.json"), "w"), ensure_ascii=False)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
