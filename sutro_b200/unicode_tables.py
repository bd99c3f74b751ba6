"""Unicode class table for the GPU pre-tokeniser (one byte per code point):
0 other, 1 letter (general category L*), 2 number (N*), 3 White_Space.
Built from the interpreter's `unicodedata`; tests/test_tokenizer_oracle.py checks it
against the `regex` module's \\p{L} / \\p{N} / \\s on every code point."""
from __future__ import annotations

import functools
import unicodedata

import numpy as np

# Unicode White_Space property (what \s means in the tokenizer regex engines).
_WHITE_SPACE = ([0x09, 0x0A, 0x0B, 0x0C, 0x0D, 0x20, 0x85, 0xA0, 0x1680, 0x2028, 0x2029, 0x202F,
                 0x205F, 0x3000] + list(range(0x2000, 0x200B)))


@functools.lru_cache(maxsize=1)
def class_table() -> np.ndarray:
    t = np.zeros(0x110000, dtype=np.uint8)
    cat = unicodedata.category
    for cp in range(0x110000):
        c = cat(chr(cp))[0]
        if c == "L":
            t[cp] = 1
        elif c == "N":
            t[cp] = 2
    t[_WHITE_SPACE] = 3
    return t
