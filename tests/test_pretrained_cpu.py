"""Loading real checkpoints (sutro_b200/pretrained.py), checked on CPU with files written by
the Hugging Face libraries themselves: `tokenizers` saves a tokenizer.json whose byte tokens
sit in GPT-2 order (not byte order), `safetensors` saves the weights, and the loader's output is
compared with what those libraries read back."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.bpe_ref import RefTokenizer
from sutro_b200 import modelspec as MS, pretrained as PT, vocab as VB

TEXTS = ["Hello, world!", "  leading spaces and 12345 numbers", "naïve café — “quotes”…",
         "日本語のテキスト 🙂", "tabs\tand\nnewlines\r\n", "it's the model's 1st re-run", ""]


def gpt2_ordered_tokenizer(v: VB.Vocab):
    """A `tokenizers.Tokenizer` over the same merges whose ids follow the GPT-2 convention:
    byte tokens 0..255 in bytes_to_unicode order, merge results in rank order after them."""
    from tokenizers import AddedToken, Regex, Tokenizer, decoders, models, pre_tokenizers
    b2u = VB._bytes_to_unicode()
    order = list(b2u.keys())                                   # GPT-2's byte order

    def s(b: bytes) -> str:
        return "".join(b2u[x] for x in b)

    vocab = {b2u[b]: i for i, b in enumerate(order)}
    merges = []
    for a, b in v.merges:
        sa, sb = s(v.token_bytes[a]), s(v.token_bytes[b])
        if sa + sb not in vocab:
            vocab[sa + sb] = len(vocab)
        merges.append((sa, sb))
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges, fuse_unk=False, byte_fallback=False))
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(VB.PRETOK_PATTERN % v.digits), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False)])
    tok.decoder = decoders.ByteLevel()
    names = [n for n, _ in sorted(v.specials.items(), key=lambda kv: kv[1])][:5]
    tok.add_special_tokens([AddedToken(n, special=True, normalized=False) for n in names])
    return tok


@pytest.mark.parametrize("family", ["qwen3", "llama"])
def test_tokenizer_json_round_trip_in_gpt2_byte_order(tmp_path, family):
    v = VB.build_vocab(family, 2048, seed=0, n_trained=600)
    hf = gpt2_ordered_tokenizer(v)
    path = str(tmp_path / "tokenizer.json")
    hf.save(path)
    loaded = PT.load_tokenizer_json(path, family, vocab_size=hf.get_vocab_size() + 11)
    assert loaded.vocab_size == hf.get_vocab_size() + 11
    assert [loaded.token_bytes[b] for b in range(256)] == [bytes([b]) for b in range(256)]
    assert sorted(loaded.id_map[:256].tolist()) == list(range(256))
    assert not np.array_equal(loaded.id_map[:256], np.arange(256))      # GPT-2 order != byte order
    assert np.array_equal(loaded.id_map[256:], np.arange(256, loaded.vocab_size))
    assert len(loaded.merges) == len(loaded.merged_ids) > 500
    ref = RefTokenizer(loaded)                       # the oracle the GPU tokenizer is tested against
    for text in TEXTS:
        engine_ids = ref.encode(text)
        assert loaded.to_real_ids(engine_ids) == hf.encode(text, add_special_tokens=False).ids
        assert loaded.decode(engine_ids).decode("utf-8") == text
    eos = "<|im_end|>" if family == "qwen3" else "<|eot_id|>"
    assert loaded.to_real_ids([loaded.eos_id]) == [hf.token_to_id(eos)]


def test_tokenizer_json_rejects_other_layouts(tmp_path):
    p = tmp_path / "t.json"
    p.write_text(json.dumps({"model": {"type": "Unigram", "vocab": []}}))
    with pytest.raises(ValueError):
        PT.load_tokenizer_json(str(p), "qwen3")
    p.write_text(json.dumps({"model": {"type": "BPE", "vocab": {"a": 0, "b": 1}, "merges": []},
                             "added_tokens": []}))
    with pytest.raises(ValueError):
        PT.load_tokenizer_json(str(p), "qwen3")      # the 256 byte tokens are not there


def hf_config(spec: MS.ModelSpec):
    cfg = {"model_type": "qwen3" if spec.family == "qwen3" else "llama",
           "hidden_size": spec.d_model, "num_attention_heads": spec.n_q_heads,
           "num_key_value_heads": spec.n_kv_heads, "head_dim": spec.head_dim,
           "num_hidden_layers": spec.n_layers, "intermediate_size": spec.d_ff,
           "vocab_size": spec.vocab_size, "tie_word_embeddings": spec.tied_embeddings,
           "rms_norm_eps": spec.rms_eps, "rope_theta": spec.rope_theta,
           "max_position_embeddings": 40960}
    if spec.rope_scaling:
        cfg["rope_scaling"] = dict(spec.rope_scaling, rope_type="llama3")
    return cfg


@pytest.mark.parametrize("name", ["tiny-qwen3", "tiny-llama", "qwen-3-4b", "llama-3.1-8b"])
def test_spec_from_hf_config(name):
    spec = MS.get_spec(name)
    got = PT.spec_from_hf_config(hf_config(spec), name, max_position=spec.max_position)
    assert got == spec
    # the same numbers transformers' config classes carry for the real models
    if name == "qwen-3-4b":
        from transformers import Qwen3Config
        c = Qwen3Config(hidden_size=2560, num_attention_heads=32, num_key_value_heads=8,
                        num_hidden_layers=36, intermediate_size=9728, vocab_size=151936,
                        tie_word_embeddings=True, rope_theta=1e6, head_dim=128)
        assert PT.spec_from_hf_config(c.to_dict(), name).n_params() == spec.n_params()


@pytest.mark.parametrize("bad", [{"model_type": "gemma3"}, {"head_dim": 64},
                                 {"num_key_value_heads": 3, "num_attention_heads": 9},
                                 {"attention_bias": True},
                                 {"rope_scaling": {"rope_type": "yarn", "factor": 4}}])
def test_unsupported_architectures_are_refused(bad):
    cfg = dict(hf_config(MS.get_spec("tiny-qwen3")), **bad)
    with pytest.raises(ValueError):
        PT.spec_from_hf_config(cfg, "x")


def test_safetensors_checkpoint_and_row_permutation(tmp_path):
    from safetensors.torch import save_file
    spec = MS.get_spec("tiny-llama")
    w = MS.make_weights(spec, seed=5)
    names = sorted(w)
    half = len(names) // 2
    save_file({k: w[k] for k in names[:half]}, str(tmp_path / "model-00001-of-00002.safetensors"))
    save_file({k: w[k] for k in names[half:]}, str(tmp_path / "model-00002-of-00002.safetensors"))
    (tmp_path / "model.safetensors.index.json").write_text(json.dumps({"weight_map": {
        k: ("model-00001-of-00002.safetensors" if i < half else "model-00002-of-00002.safetensors")
        for i, k in enumerate(names)}}))
    got = PT.load_hf_weights(str(tmp_path))
    assert sorted(got) == names and all(torch.equal(got[k], w[k]) for k in names)
    perm = np.arange(spec.vocab_size, dtype=np.int32)
    perm[:256] = np.random.RandomState(0).permutation(256)
    ew = MS.pack_for_engine(spec, got, "cpu", row_map=perm)
    assert torch.equal(ew.embed[7], w["model.embed_tokens.weight"][perm[7]])
    assert torch.equal(ew.lm_head[300], w["lm_head.weight"][300])
    plain = MS.pack_for_engine(spec, got, "cpu")
    assert torch.equal(plain.embed, w["model.embed_tokens.weight"])
    with pytest.raises(FileNotFoundError):
        PT.load_hf_weights(str(tmp_path / "nowhere"))


def test_real_pretokenizer_patterns_and_normalizers(tmp_path):
    """The Split regex as Qwen2/3 and Llama-3 files spell it, NFC normalizer, and refusals."""
    v = VB.build_vocab("qwen3", 2048, seed=0, n_trained=600)
    hf = gpt2_ordered_tokenizer(v)
    path = str(tmp_path / "tokenizer.json")
    hf.save(path)
    tj = json.load(open(path))
    qwen_rx = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*"
               r"|\s*[\r\n]+|\s+(?!\S)|\s+")
    tj["pre_tokenizer"]["pretokenizers"][0]["pattern"] = {"Regex": qwen_rx}
    tj["normalizer"] = {"type": "NFC"}
    json.dump(tj, open(path, "w"))
    loaded = PT.load_tokenizer_json(path, "qwen3")
    assert loaded.digits == 1 and loaded.normalize_nfc
    tj["pre_tokenizer"]["pretokenizers"][0]["pattern"] = {"Regex": qwen_rx.replace(r"\p{N}|", r"\p{N}{1,3}|")}
    tj["normalizer"] = None
    json.dump(tj, open(path, "w"))
    assert PT.load_tokenizer_json(path, "qwen3").digits == 3
    tj["pre_tokenizer"]["pretokenizers"][0]["pattern"] = {"Regex": r"\w+|\s+"}
    json.dump(tj, open(path, "w"))
    with pytest.raises(ValueError):
        PT.load_tokenizer_json(path, "qwen3")
    tj["pre_tokenizer"] = None
    tj["normalizer"] = {"type": "Lowercase"}
    json.dump(tj, open(path, "w"))
    with pytest.raises(ValueError):
        PT.load_tokenizer_json(path, "qwen3")
    from sutro_b200.engine import _nfc_rows
    assert _nfc_rows(["é", None, "plain", 3]) == ["é", None, "plain", 3]


def test_ignore_merges_is_honoured_by_proof_or_refused(tmp_path, monkeypatch):
    """Llama-3 tokenizer files set model.ignore_merges (a pre-token that is itself a vocabulary
    entry is emitted whole).  The GPU BPE always merges, so the loader PROVES the flag is a no-op
    — merging every vocabulary string must reproduce its own token — and otherwise refuses to
    load (or loads with an explicit opt-in), never serving other ids silently."""
    v = VB.build_vocab("llama", 2048, seed=0, n_trained=600)
    hf = gpt2_ordered_tokenizer(v)
    path = str(tmp_path / "tokenizer.json")
    hf.save(path)
    tj = json.load(open(path))
    assert PT.load_tokenizer_json(path, "llama").vocab_size > 0      # flag absent: loads
    tj["model"]["ignore_merges"] = True
    json.dump(tj, open(path, "w"))
    # does merging reproduce every entry of this vocabulary?  (count with the oracle's BPE)
    monkeypatch.setenv("SB200_ALLOW_IGNORE_MERGES_MISMATCH", "1")
    loaded = PT.load_tokenizer_json(path, "llama")
    ref = RefTokenizer(loaded)
    mismatches = sum(1 for e, tb in enumerate(loaded.token_bytes)
                     if len(tb) >= 2 and ref._bpe(tb) != [e])
    monkeypatch.delenv("SB200_ALLOW_IGNORE_MERGES_MISMATCH")
    if mismatches:
        with pytest.raises(ValueError, match=f"ignore_merges is true and {mismatches} vocabulary"):
            PT.load_tokenizer_json(path, "llama")
    else:
        assert PT.load_tokenizer_json(path, "llama").vocab_size == loaded.vocab_size
    # a vocabulary reduced to its byte tokens and first merges is consistent: accepted
    keep = dict(list(tj["model"]["vocab"].items())[:256 + 40])
    tj["model"]["vocab"] = keep
    tj["model"]["merges"] = [m for m in tj["model"]["merges"]
                             if "".join(m if isinstance(m, list) else m.split(" ")) in keep][:40]
    tj["added_tokens"] = [dict(a, id=len(keep) + i) for i, a in enumerate(tj["added_tokens"])]
    json.dump(tj, open(path, "w"))
    small = PT.load_tokenizer_json(path, "llama")
    assert len(small.merges) > 0
