"""Pins the tokenizer oracle (oracle/bpe_ref.py) and the Unicode class table against the
third-party libraries that define the algorithm: `tokenizers` and `regex`."""
import numpy as np
import pytest
import regex

from oracle.bpe_ref import RefTokenizer
from sutro_b200 import synth, vocab as V
from sutro_b200.unicode_tables import class_table

TEXTS = synth.product_reviews(60, seed=5) + synth.extraction_documents(10, seed=6) + [
    "", " ", "  ", "a", "don't DON'T we'LL they've I'm he'd it's 'tis",
    "tabs\tand\nnewlines\r\n\r\n  indented   \n\n\n", "trailing spaces   ", "   leading",
    "numbers 1234567 3.14159 1,000,000 v2.1", "punct!!! ... ((nested)) [x]{y} #tag @user",
    "unicode: café naïve Straße 東京 Привет мир ١٢٣ ½ 𝒳 🙂🙂 end", "a b c　d",
    "aaaaaaaaaaaaaaaa bbbbbbbb abababababab", "x" * 300, "mixed123abc456 7z", "\n", "\n\n a",
    "'", "''s", " 's", "it 's", "O'Re 'LLama", "end with quote'",
]


@pytest.mark.parametrize("family,size", [("qwen3", 2048), ("llama", 1024)])
def test_ref_tokenizer_matches_hf_tokenizers(family, size):
    v = V.build_vocab(family, size, seed=0, n_trained=600)
    ref = RefTokenizer(v)
    hf = V.to_hf_tokenizer(v)
    for t in TEXTS:
        assert ref.encode(t) == hf.encode(t, add_special_tokens=False).ids, repr(t)
        assert v.decode(ref.encode(t)) == t.encode("utf-8")


def test_full_size_vocab_matches_hf_on_samples():
    v = V.build_vocab("qwen3", 151936, seed=0)
    assert v.specials["<|endoftext|>"] == 151643 and v.specials["<|im_end|>"] == 151645
    ref = RefTokenizer(v)
    hf = V.to_hf_tokenizer(v)
    for t in TEXTS[:25] + TEXTS[-12:]:
        assert ref.encode(t) == hf.encode(t, add_special_tokens=False).ids, repr(t)
    n_tok = sum(len(ref.encode(t)) for t in TEXTS[:40])
    n_byte = sum(len(t.encode()) for t in TEXTS[:40])
    assert n_byte / n_tok > 3.0  # trained merges compress ordinary text


def test_template_rendering_matches_hf_with_special_tokens():
    v = V.build_vocab("qwen3", 2048, seed=0, n_trained=600)
    ref = RefTokenizer(v)
    hf = V.to_hf_tokenizer(v)
    tpl = V.chat_template("qwen3", synth.README_SYSTEM_PROMPT)
    for row in synth.README_REVIEWS:
        text = "".join(tpl.prefix) + row + "".join(tpl.suffix)
        assert ref.render(tpl, row) == hf.encode(text, add_special_tokens=False).ids


def test_unicode_class_table_matches_regex_module():
    t = class_table()
    pl, pn, ps = regex.compile(r"\p{L}"), regex.compile(r"\p{N}"), regex.compile(r"\s")
    rng = np.random.RandomState(0)
    cps = list(range(0x3000)) + rng.randint(0x3000, 0x110000, size=20000).tolist()
    import unicodedata
    bad = 0
    for cp in cps:
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        want = 1 if pl.match(ch) else 2 if pn.match(ch) else 3 if ps.match(ch) else 0
        if want == 3 and 0x1C <= cp <= 0x1F:
            continue  # `regex` counts the ASCII separators as \s; the tokenizer engines do not
        if t[cp] != want and unicodedata.category(ch) == "Cn":
            continue  # assigned after the interpreter's Unicode version: library version skew
        bad += int(t[cp] != want)
    assert bad == 0, bad
