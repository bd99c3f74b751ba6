"""Deterministic synthetic byte-level BPE vocabularies and chat templates.

No tokenizer files exist offline, so the engine ships a seeded
generator that produces a *real* byte-level BPE (256 byte tokens + ordered merges +
special tokens) of the model's vocabulary size:
  1. merges learned by plain BPE training on a synthetic English-like corpus
     (synth.py), so that ordinary text compresses to ~4 bytes/token;
  2. filler merges — seeded random pairs of existing tokens with distinct byte
     strings — up to the model's regular-token count, so every id the lm_head can
     emit decodes to bytes;
  3. special tokens laid out like the model family's (Qwen: <|endoftext|>,
     <|im_start|>, <|im_end|> right after the regular tokens).
Token id == rank order: ids 0..255 are the raw bytes, id 256+i is merge i.

The reference repo is silent on templating (the server does it, sutro/sdk.py:196-208
just forwards `system_prompt`).  Builder decision, stated in every report: ChatML for
Qwen3 (no thinking block), the Llama-3 header format for Llama.
"""
from __future__ import annotations

import functools
from collections import Counter
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

# Pre-tokenisation follows the published GPT-4-style pattern used by Qwen2/3 and
# Llama-3 tokenizers.  digits=1 for Qwen (\p{N}), 3 for Llama-3 (\p{N}{1,3}).
PRETOK_PATTERN = (r"(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}{1,%d}|"
                  r" ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+|\s+(?!\S)|\s+")


@dataclass
class Vocab:
    family: str
    vocab_size: int
    token_bytes: List[bytes]            # per id (b"" for special / unused ids)
    merges: List[Tuple[int, int]]       # merge i: (left id, right id) -> id 256+i
    specials: Dict[str, int]
    digits: int                         # max digits per pre-token
    # real tokenizer files (pretrained.load_tokenizer_json): id of merge i's result when it is
    # not 256+i, and the engine-id -> real-id permutation of the 256 byte tokens
    merged_ids: Optional[List[int]] = None
    id_map: Optional[np.ndarray] = None
    normalize_nfc: bool = False         # the tokenizer file asks for NFC-normalised input
    # `ignore_merges` of the tokenizer file (Llama-3): a pre-token that is itself a vocabulary
    # entry is emitted as that token without merging.  None = flag off.  When on: the entries
    # whose merges do not rebuild them, as (the sequence their merges do build, entry id) —
    # what the GPU tokenizer needs to honour the flag (sb200_tokenizer_set_word_overrides).
    word_overrides: Optional[List[Tuple[Tuple[int, ...], int]]] = None

    def to_real_ids(self, ids):
        """engine token ids -> the tokenizer file's ids (identity for synthetic vocabularies)"""
        return list(ids) if self.id_map is None else [int(self.id_map[i]) for i in ids]

    @property
    def eos_id(self) -> int:
        return self.specials["<|im_end|>" if self.family == "qwen3" else "<|eot_id|>"]

    @property
    def n_regular(self) -> int:
        return 256 + len(self.merges)

    # ---- flat arrays for the engine ------------------------------------
    def byte_blob(self) -> Tuple[np.ndarray, np.ndarray]:
        off = np.zeros(self.vocab_size + 1, dtype=np.int32)
        lens = np.fromiter((len(b) for b in self.token_bytes), dtype=np.int64,
                           count=self.vocab_size)
        off[1:] = np.cumsum(lens)
        blob = np.frombuffer(b"".join(self.token_bytes), dtype=np.uint8).copy()
        return blob, off

    def merge_array(self) -> np.ndarray:
        return np.asarray(self.merges, dtype=np.int32).reshape(-1, 2)

    def decode(self, ids) -> bytes:
        return b"".join(self.token_bytes[i] for i in ids)

    def decode_text(self, ids) -> str:
        return self.decode(ids).decode("utf-8", errors="replace")


# --------------------------------------------------------------------------- training
def _train_merges(corpus_words: Counter, n_merges: int) -> List[Tuple[bytes, bytes]]:
    """Classic BPE training over a word-frequency table (deterministic: ties broken by
    the pair's byte strings)."""
    words = {w: [bytes([b]) for b in w] for w in corpus_words}
    merges: List[Tuple[bytes, bytes]] = []
    for _ in range(n_merges):
        pairs: Counter = Counter()
        for w, syms in words.items():
            f = corpus_words[w]
            for a, b in zip(syms, syms[1:]):
                pairs[(a, b)] += f
        if not pairs:
            break
        best = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        if best[1] < 2:
            break
        a, b = best[0]
        merges.append((a, b))
        ab = a + b
        for w, syms in words.items():
            if len(syms) < 2:
                continue
            out, i = [], 0
            while i < len(syms):
                if i + 1 < len(syms) and syms[i] == a and syms[i + 1] == b:
                    out.append(ab)
                    i += 2
                else:
                    out.append(syms[i])
                    i += 1
            words[w] = out
    return merges


@functools.lru_cache(maxsize=8)
def build_vocab(family: str, vocab_size: int, seed: int = 0, n_trained: int = 3000) -> Vocab:
    import regex
    from . import synth

    digits = 1 if family == "qwen3" else 3
    n_special = 293 if vocab_size >= 100000 else 16
    n_regular = vocab_size - n_special
    assert n_regular > 300, "vocabulary too small"

    pat = regex.compile(PRETOK_PATTERN % digits)
    text = "\n".join(synth.corpus_sentences(4000, seed=1234))
    words = Counter(m.group(0).encode("utf-8") for m in pat.finditer(text))
    trained = _train_merges(words, min(n_trained, n_regular - 256))

    tok2id: Dict[bytes, int] = {bytes([i]): i for i in range(256)}
    token_bytes: List[bytes] = [bytes([i]) for i in range(256)]
    merges: List[Tuple[int, int]] = []
    for a, b in trained:
        ab = a + b
        if ab in tok2id:
            continue
        merges.append((tok2id[a], tok2id[b]))
        tok2id[ab] = len(token_bytes)
        token_bytes.append(ab)

    rng = np.random.RandomState(seed)
    while len(token_bytes) < n_regular:
        cur = len(token_bytes)
        cand = rng.randint(0, cur, size=(4096, 2))
        for a, b in cand:
            ab = token_bytes[a] + token_bytes[b]
            if len(ab) > 12 or ab in tok2id:
                continue
            merges.append((int(a), int(b)))
            tok2id[ab] = len(token_bytes)
            token_bytes.append(ab)
            if len(token_bytes) == n_regular:
                break

    if family == "qwen3":
        names = ["<|endoftext|>", "<|im_start|>", "<|im_end|>"]
    else:
        names = ["<|begin_of_text|>", "<|end_of_text|>", "<|start_header_id|>",
                 "<|end_header_id|>", "<|eot_id|>"]
    specials: Dict[str, int] = {}
    for i in range(n_special):
        name = names[i] if i < len(names) else f"<|extra_{i}|>"
        specials[name] = n_regular + i
        token_bytes.append(b"")
    assert len(token_bytes) == vocab_size
    return Vocab(family, vocab_size, token_bytes, merges, specials, digits)


# --------------------------------------------------------------------------- templates
@dataclass
class Template:
    """A prompt is  prefix_pieces + row text + suffix_pieces ; pieces are either special
    token names or plain text (tokenised independently, as HF does around specials)."""
    prefix: List[str]
    suffix: List[str]


def chat_template(family: str, system_prompt: str | None,
                  empty_think_block: bool = False) -> Template:
    """`empty_think_block`: Qwen3's released template in non-thinking mode opens the answer with
    an empty `<think>` block; real Qwen3 checkpoints are served that way (pretrained.py), the
    synthetic benchmark models are not (there is nothing to suppress)."""
    if family == "qwen3":
        pre: List[str] = []
        if system_prompt:
            pre += ["<|im_start|>", "system\n" + system_prompt, "<|im_end|>", "\n"]
        pre += ["<|im_start|>", "user\n"]
        suf = ["<|im_end|>", "\n", "<|im_start|>", "assistant\n"]
        if empty_think_block:
            # <think> / </think> are added tokens of the Qwen3 vocabulary: separate pieces, so
            # that they map to their own ids when the tokenizer file defines them
            suf += ["<think>", "\n\n", "</think>", "\n\n"]
        return Template(pre, suf)
    pre = ["<|begin_of_text|>"]
    if system_prompt:
        pre += ["<|start_header_id|>", "system", "<|end_header_id|>", "\n\n" + system_prompt,
                "<|eot_id|>"]
    pre += ["<|start_header_id|>", "user", "<|end_header_id|>", "\n\n"]
    suf = ["<|eot_id|>", "<|start_header_id|>", "assistant", "<|end_header_id|>", "\n\n"]
    return Template(pre, suf)


def embedding_template(family: str) -> Template:
    # Qwen3-Embedding appends <|endoftext|> and pools its hidden state.
    return Template([], ["<|endoftext|>" if family == "qwen3" else "<|end_of_text|>"])


# --------------------------------------------------------------------------- HF interop (oracle side)
def _bytes_to_unicode() -> Dict[int, str]:
    bs = (list(range(ord("!"), ord("~") + 1)) + list(range(ord("¡"), ord("¬") + 1)) +
          list(range(ord("®"), ord("ÿ") + 1)))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return {b: chr(c) for b, c in zip(bs, cs)}


def to_hf_tokenizer(v: Vocab):
    """Build a `tokenizers.Tokenizer` with the identical vocabulary — used by the
    tokenizer ORACLE (tests), never by the product path."""
    from tokenizers import Regex, Tokenizer, decoders, models, pre_tokenizers
    b2u = _bytes_to_unicode()

    def s(b: bytes) -> str:
        return "".join(b2u[x] for x in b)

    vocab = {s(tb): i for i, tb in enumerate(v.token_bytes[:v.n_regular])}
    merges = [(s(v.token_bytes[a]), s(v.token_bytes[b])) for a, b in v.merges]
    tok = Tokenizer(models.BPE(vocab=vocab, merges=merges, fuse_unk=False, byte_fallback=False))
    tok.pre_tokenizer = pre_tokenizers.Sequence([
        pre_tokenizers.Split(Regex(PRETOK_PATTERN % v.digits), behavior="isolated", invert=False),
        pre_tokenizers.ByteLevel(add_prefix_space=False, use_regex=False),
    ])
    tok.decoder = decoders.ByteLevel()
    from tokenizers import AddedToken
    tok.add_special_tokens([AddedToken(n, special=True, normalized=False)
                            for n, _ in sorted(v.specials.items(), key=lambda kv: kv[1])])
    for n, i in v.specials.items():
        assert tok.token_to_id(n) == i, (n, i, tok.token_to_id(n))
    return tok
