"""Seeded synthetic text for the BASELINE.json configurations (no datasets offline:
the README's hf://datasets/sutro/synthetic-product-reviews-20k is unreachable).

Every generator is a pure function of its seed so the engine, the oracle and the
benchmark see identical bytes.
"""
from __future__ import annotations

from typing import List

import numpy as np

_PRODUCTS = ["headphones", "blender", "backpack", "keyboard", "water bottle", "desk lamp",
             "running shoes", "coffee maker", "phone case", "monitor", "office chair", "tent",
             "air fryer", "smart watch", "vacuum cleaner", "toaster", "bike helmet", "speaker",
             "mattress", "router", "camera", "drill", "jacket", "sunglasses", "notebook"]
_ADJ_POS = ["excellent", "fantastic", "reliable", "sturdy", "comfortable", "impressive",
            "lightweight", "beautiful", "quiet", "durable", "intuitive", "responsive"]
_ADJ_NEG = ["disappointing", "flimsy", "noisy", "uncomfortable", "overpriced", "unreliable",
            "cheap-feeling", "slow", "bulky", "defective", "confusing", "fragile"]
_ADJ_NEU = ["average", "decent", "acceptable", "ordinary", "fine", "standard", "unremarkable"]
_OPENERS = ["I bought this {p} last month and", "After two weeks with the {p},",
            "My {p} arrived yesterday;", "Honestly, this {p}", "We ordered the {p} for our home and",
            "As a long-time user of similar products, this {p}", "The {p}", "For the price, the {p}"]
_MIDDLES = ["it is {a}", "it feels {a} in daily use", "the build quality is {a}",
            "I found it {a} compared to my old one", "the overall experience has been {a}",
            "everyone in my family agrees it is {a}", "setup was {a}", "the battery life is {a}"]
_CLOSERS = ["I would recommend it to a friend.", "I will probably return it.",
            "It does what it says, nothing more.", "Five stars from me!", "Two stars at best.",
            "Shipping took 3 days and the box was intact.", "Customer support answered in 24 hours.",
            "I paid $49.99 and it was worth every cent.", "Don't waste your money.",
            "It's okay, I guess.", "Can't complain so far.", "Would buy again in 2025."]
_EXTRAS = ["The color matches the photos.", "Instructions were only in English and German.",
           "It fits perfectly on my desk.", "The strap broke after 10 uses.",
           "My kids love it, too.", "The app needs an update (v2.1 crashes).",
           "Packaging was 100% recyclable.", "It weighs about 1.5 kg.",
           "I've had no issues whatsoever.", "The warranty covers 12 months."]
_DOC_WORDS = ("the of and to in is that for it as was with be by on not he this are or his from at "
              "which but have an had they you were their one all we can her has there been if more "
              "when will would who so no out up said what its about than into them only other new "
              "some could time these two may then do first any my now such like our over man me even "
              "most made after also did many before must through back years where much your way well "
              "down should because each just those people how too little state good very make world "
              "still own see men work long get here between both life being under never day same "
              "another know while last might us great old year off come since against go came right "
              "used take three system data model report analysis customer product service quality "
              "price delivery order account invoice payment contract manager engineer meeting").split()
_NAMES = ["Alice", "Bob", "Carla", "Deepak", "Elena", "Farid", "Grace", "Hiro", "Ines", "Jonas",
          "Keiko", "Liam", "Maya", "Noor", "Oscar", "Priya", "Quinn", "Rosa", "Sven", "Tara"]
_CITIES = ["Berlin", "Austin", "Osaka", "Lagos", "Lima", "Oslo", "Pune", "Lyon", "Perth", "Quito"]


def _review(rng: np.random.RandomState, target_sentences: int) -> str:
    tone = rng.randint(3)
    adjs = (_ADJ_POS, _ADJ_NEU, _ADJ_NEG)[tone]
    p = _PRODUCTS[rng.randint(len(_PRODUCTS))]
    parts = [_OPENERS[rng.randint(len(_OPENERS))].format(p=p) + " " +
             _MIDDLES[rng.randint(len(_MIDDLES))].format(a=adjs[rng.randint(len(adjs))]) + "."]
    for _ in range(target_sentences - 1):
        k = rng.randint(3)
        if k == 0:
            parts.append("Also, " + _MIDDLES[rng.randint(len(_MIDDLES))].format(
                a=adjs[rng.randint(len(adjs))]) + ".")
        elif k == 1:
            parts.append(_EXTRAS[rng.randint(len(_EXTRAS))])
        else:
            parts.append(_CLOSERS[rng.randint(len(_CLOSERS))])
    return " ".join(parts)


def corpus_sentences(n: int, seed: int) -> List[str]:
    """Training corpus for the synthetic BPE merges (vocab.py)."""
    rng = np.random.RandomState(seed)
    out = [_review(rng, 1 + rng.randint(5)) for _ in range(n // 2)]
    for _ in range(n - len(out)):
        k = 8 + rng.randint(24)
        out.append(" ".join(_DOC_WORDS[i] for i in rng.randint(0, len(_DOC_WORDS), size=k)) + ".")
    return out


def product_reviews(n: int, seed: int = 0) -> List[str]:
    """Config 2: product reviews, length ~ lognormal, clipped to roughly
    16-256 tokens, mean ~ 96 tokens under the synthetic BPE vocabulary (a sentence of the
    phrase grammar is ~9.4 tokens; measured by tests/test_bench_cpu.py)."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        sents = int(np.clip(np.round(rng.lognormal(mean=2.2, sigma=0.5)), 2, 27))
        out.append(_review(rng, sents))
    return out


def documents(n: int, seed: int = 1, words: int = 380) -> List[str]:
    """Config 3: long pseudo-documents (token length is fixed up by the caller through
    truncation/padding at the token level)."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        idx = rng.randint(0, len(_DOC_WORDS), size=words)
        toks = [_DOC_WORDS[i] for i in idx]
        for j in range(12, words, 17):
            toks[j] = toks[j] + "."
        out.append(" ".join(toks))
    return out


def short_texts(n: int, seed: int = 2) -> List[str]:
    """Config 4: short texts, 8-64 tokens."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        k = 4 + rng.randint(40)
        out.append(" ".join(_DOC_WORDS[i] for i in rng.randint(0, len(_DOC_WORDS), size=k)))
    return out


def extraction_documents(n: int, seed: int = 3) -> List[str]:
    """Config 5: pseudo-documents mentioning people, cities, amounts."""
    rng = np.random.RandomState(seed)
    out = []
    for _ in range(n):
        k = 2 + rng.randint(5)
        parts = []
        for _ in range(k):
            parts.append(
                f"{_NAMES[rng.randint(len(_NAMES))]} from {_CITIES[rng.randint(len(_CITIES))]} "
                f"ordered {1 + rng.randint(9)} {_PRODUCTS[rng.randint(len(_PRODUCTS))]} for "
                f"${10 + rng.randint(990)}.{rng.randint(100):02d} on 2025-{1 + rng.randint(12):02d}-"
                f"{1 + rng.randint(28):02d}.")
            filler = 10 + rng.randint(40)
            parts.append(" ".join(_DOC_WORDS[i] for i in
                                  rng.randint(0, len(_DOC_WORDS), size=filler)) + ".")
        out.append(" ".join(parts))
    return out


README_REVIEWS = [
    "The battery life is terrible.",
    "Great camera and build quality!",
    "Too expensive for what it offers.",
]
README_SYSTEM_PROMPT = "Classify the sentiment of the review as positive, neutral, or negative."
