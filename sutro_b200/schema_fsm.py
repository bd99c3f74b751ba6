"""output_schema -> byte-level DFA for constrained decoding (host side of K8).

The reference only normalises the schema (sutro/common.py:152-163: a Pydantic class
becomes `.model_json_schema()`, a dict passes through) and ships it as `json_schema`
in the request (sutro/sdk.py:199); enforcing it is the server's job.  Locally the
schema is compiled to a deterministic automaton over UTF-8 bytes that accepts exactly
the *compact* JSON serialisations (no optional white space, properties in declaration
order) of instances of the schema; the GPU turns it into per-state token masks
(csrc/sampler_fsm.cu).

Supported subset (what Pydantic emits): object/properties (all listed properties are
emitted, in order), dict-like objects (additionalProperties / propertyNames /
min-/maxProperties), string (minLength/maxLength, pattern, format date / date-time / time /
uuid / email / ipv4 / uri, enum, const), integer / number (minimum / maximum / exclusive
bounds — exact digit automata — and multipleOf over small ranges), boolean, null, enum of
JSON literals, array (items, prefixItems, minItems/maxItems, uniqueItems over small enums),
anyOf/oneOf, allOf of compatible parts, $ref/$defs (recursion unrolled to a fixed depth).  A keyword that would
constrain the output but is not implemented raises SchemaError — nothing is silently
ignored.  Unbounded strings / arrays / digits / repeats get explicit caps (`FsmLimits`) so
that every path through the automaton terminates — with random weights a model never chooses
to stop on its own.  Soundness (every accepted string validates) is tested against pydantic
and a small validator; the token masks are cross-checked against xgrammar.
"""
from __future__ import annotations

import itertools
import json
from dataclasses import dataclass
from decimal import Decimal
from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class FsmLimits:
    max_string_chars: int = 64     # cap when a string has no maxLength
    max_array_items: int = 8       # cap when an array has no maxItems
    max_int_digits: int = 9        # digits of an unbounded integer
    max_frac_digits: int = 4
    small_int_range: int = 2048    # ranges up to this size are encoded exactly
    max_recursion: int = 2         # how often a recursive $ref may be re-entered


@dataclass
class ByteDFA:
    trans: np.ndarray    # int32 [n_states, 256], -1 = dead
    accept: np.ndarray   # uint8 [n_states]
    final: np.ndarray    # uint8 [n_states]: accepting and no outgoing edge
    start: int = 0

    @property
    def n_states(self) -> int:
        return self.trans.shape[0]

    def matches(self, data: bytes) -> bool:
        s = self.start
        for b in data:
            s = int(self.trans[s, b])
            if s < 0:
                return False
        return bool(self.accept[s])


    def longest_path(self) -> Optional[int]:
        """Length in bytes of the longest string the automaton accepts, or None when it has a
        cycle (an unbounded language).  Every token is at least one byte, so this bounds the
        number of tokens a constrained row can need: the SDK sizes max_new_tokens from it."""
        n = self.n_states
        succ = [np.unique(self.trans[s][self.trans[s] >= 0]) for s in range(n)]
        depth = [-1] * n          # longest path to a dead end, from s
        state = [0] * n           # 0 new, 1 on stack, 2 done
        stack = [(self.start, 0)]
        while stack:
            s, i = stack.pop()
            if i == 0:
                if state[s] == 2:
                    continue
                state[s] = 1
            if i < len(succ[s]):
                stack.append((s, i + 1))
                t = int(succ[s][i])
                if state[t] == 1:
                    return None
                if state[t] == 0:
                    stack.append((t, 0))
            else:
                depth[s] = max([1 + depth[int(t)] for t in succ[s]], default=0)
                state[s] = 2
        return depth[self.start]

    # ---- forced runs (jump-forward decoding) --------------------------------------
    def forced_run(self, state: int):
        """Follow `state` while exactly one byte keeps the automaton alive and the state is
        not accepting (an accepting state is a decision: stop or continue).  Returns
        (bytes, end_state)."""
        out = bytearray()
        s = state
        while not self.accept[s]:
            nxt = np.nonzero(self.trans[s] >= 0)[0]
            if len(nxt) != 1:
                break
            out.append(int(nxt[0]))
            s = int(self.trans[s, nxt[0]])
        return bytes(out), s

    def forced_plan(self):
        """Jump-forward plan: the byte string every output must start with (+ the state
        after it), and for every state whose continuation is forced all the way to a final
        state, that terminal tail.  With constrained decoding the model has no say over
        these bytes, so the engine feeds the prefix as part of the prompt and appends tails
        without running the model (same strings, fewer decode steps)."""
        prefix, start_after = self.forced_run(self.start)
        tails = {}
        for s in range(self.n_states):
            run, end = self.forced_run(s)
            if run and self.final[end]:
                tails[s] = run
        return prefix, start_after, tails


# --------------------------------------------------------------------------- NFA
class _NFA:
    def __init__(self):
        self.eps: List[List[int]] = []
        self.tr: List[List[Tuple[int, int]]] = []   # (256-bit byte mask, target)

    def new(self) -> int:
        self.eps.append([])
        self.tr.append([])
        return len(self.eps) - 1


Frag = Tuple[int, int]


def _mask(*ranges) -> int:
    m = 0
    for r in ranges:
        if isinstance(r, int):
            m |= 1 << r
        else:
            lo, hi = r
            for b in range(lo, hi + 1):
                m |= 1 << b
    return m


class _Builder:
    def __init__(self):
        self.n = _NFA()

    def empty(self) -> Frag:
        s = self.n.new()
        return (s, s)

    def bset(self, mask: int) -> Frag:
        s, e = self.n.new(), self.n.new()
        self.n.tr[s].append((mask, e))
        return (s, e)

    def lit(self, data: bytes) -> Frag:
        s = cur = self.n.new()
        for b in data:
            nx = self.n.new()
            self.n.tr[cur].append((1 << b, nx))
            cur = nx
        return (s, cur)

    def seq(self, *fs: Frag) -> Frag:
        fs = [f for f in fs if f is not None]
        for a, b in zip(fs, fs[1:]):
            self.n.eps[a[1]].append(b[0])
        return (fs[0][0], fs[-1][1])

    def alt(self, *fs: Frag) -> Frag:
        s, e = self.n.new(), self.n.new()
        for f in fs:
            self.n.eps[s].append(f[0])
            self.n.eps[f[1]].append(e)
        return (s, e)

    def opt(self, f: Frag) -> Frag:
        s, e = self.n.new(), self.n.new()
        self.n.eps[s] += [f[0], e]
        self.n.eps[f[1]].append(e)
        return (s, e)

    def rep(self, make: Callable[[], Frag], lo: int, hi: int) -> Frag:
        """make(){lo,hi} with fresh copies; linear in hi."""
        s = cur = self.n.new()
        e = self.n.new()
        if lo == 0:
            self.n.eps[s].append(e)
        for i in range(1, hi + 1):
            f = make()
            self.n.eps[cur].append(f[0])
            cur = f[1]
            if i >= lo:
                self.n.eps[cur].append(e)
        return (s, e)

    def star(self, make: Callable[[], Frag]) -> Frag:
        """make()* as a real loop (no counting, so no blow-up when it is determinised)."""
        s = self.n.new()
        f = make()
        self.n.eps[s].append(f[0])
        self.n.eps[f[1]].append(s)
        return (s, s)

    def literals(self, words: List[bytes]) -> Frag:
        """alternation of byte strings as a trie (keeps the NFA small for enums/ranges)."""
        s, e = self.n.new(), self.n.new()
        trie: Dict[Tuple[int, bytes], int] = {}
        for w in words:
            cur = s
            for i, b in enumerate(w):
                key = (cur, bytes([b]))
                if key not in trie:
                    nx = self.n.new()
                    self.n.tr[cur].append((1 << b, nx))
                    trie[key] = nx
                cur = trie[key]
            self.n.eps[cur].append(e)
        return (s, e)

    # ---- JSON pieces ---------------------------------------------------
    def json_char(self) -> Frag:
        """one JSON string character (one code point, or one escape sequence)."""
        b = self
        ascii_ok = _mask((0x20, 0x21), (0x23, 0x5B), (0x5D, 0x7F))
        cont = _mask((0x80, 0xBF))
        two = b.seq(b.bset(_mask((0xC2, 0xDF))), b.bset(cont))
        three = b.alt(
            b.seq(b.bset(_mask(0xE0)), b.bset(_mask((0xA0, 0xBF))), b.bset(cont)),
            b.seq(b.bset(_mask((0xE1, 0xEC), (0xEE, 0xEF))), b.bset(cont), b.bset(cont)),
            b.seq(b.bset(_mask(0xED)), b.bset(_mask((0x80, 0x9F))), b.bset(cont)))
        four = b.alt(
            b.seq(b.bset(_mask(0xF0)), b.bset(_mask((0x90, 0xBF))), b.bset(cont), b.bset(cont)),
            b.seq(b.bset(_mask((0xF1, 0xF3))), b.bset(cont), b.bset(cont), b.bset(cont)),
            b.seq(b.bset(_mask(0xF4)), b.bset(_mask((0x80, 0x8F))), b.bset(cont), b.bset(cont)))
        hexd = _mask((0x30, 0x39), (0x41, 0x46), (0x61, 0x66))
        esc = b.seq(b.bset(_mask(0x5C)), b.alt(
            b.bset(_mask(*[ord(c) for c in '"\\/bfnrt'])),
            # \\uXXXX except the surrogate block D800-DFFF (lone surrogates are not text)
            b.seq(b.bset(_mask(ord("u"))), b.alt(
                b.seq(b.bset(hexd & ~_mask(ord("d"), ord("D"))), b.bset(hexd), b.bset(hexd),
                      b.bset(hexd)),
                b.seq(b.bset(_mask(ord("d"), ord("D"))), b.bset(_mask((0x30, 0x37))),
                      b.bset(hexd), b.bset(hexd))))))
        return b.alt(b.bset(ascii_ok), two, three, four, esc)

    def json_string(self, lo: int, hi: int) -> Frag:
        q = _mask(0x22)
        return self.seq(self.bset(q), self.rep(self.json_char, lo, hi), self.bset(q))

    # ---- digit strings --------------------------------------------------
    def digits_between(self, x: str, y: str, dot_before: Optional[int] = None) -> Frag:
        """Equal-length digit strings d with x <= d <= y (compared as numbers).  When
        `dot_before` is given, a '.' is emitted before the digit at that index."""
        assert len(x) == len(y) and x <= y
        b = self
        any_d = _mask((0x30, 0x39))

        def digit(i: int, mask: int) -> Frag:
            f = b.bset(mask)
            return b.seq(b.lit(b"."), f) if dot_before is not None and i == dot_before else f

        def free(i: int) -> Optional[Frag]:      # any digits from index i to the end
            fs = [digit(k, any_d) for k in range(i, len(x))]
            return b.seq(*fs) if fs else None

        def at_least(i: int) -> Optional[Frag]:  # suffix >= x[i:]
            if i == len(x):
                return None
            d = int(x[i])
            alts = [b.seq(digit(i, 1 << (0x30 + d)), at_least(i + 1))]
            if d < 9:
                alts.append(b.seq(digit(i, _mask((0x30 + d + 1, 0x39))), free(i + 1)))
            return alts[0] if len(alts) == 1 else b.alt(*alts)

        def at_most(i: int) -> Optional[Frag]:   # suffix <= y[i:]
            if i == len(y):
                return None
            d = int(y[i])
            alts = [b.seq(digit(i, 1 << (0x30 + d)), at_most(i + 1))]
            if d > 0:
                alts.append(b.seq(digit(i, _mask((0x30, 0x30 + d - 1))), free(i + 1)))
            return alts[0] if len(alts) == 1 else b.alt(*alts)

        def between(i: int) -> Optional[Frag]:
            if i == len(x):
                return None
            dx, dy = int(x[i]), int(y[i])
            if dx == dy:
                return b.seq(digit(i, 1 << (0x30 + dx)), between(i + 1))
            alts = [b.seq(digit(i, 1 << (0x30 + dx)), at_least(i + 1)),
                    b.seq(digit(i, 1 << (0x30 + dy)), at_most(i + 1))]
            if dy - dx > 1:
                alts.append(b.seq(digit(i, _mask((0x30 + dx + 1, 0x30 + dy - 1))), free(i + 1)))
            return b.alt(*alts)

        return between(0)

    def scaled_range(self, lo: int, hi: int, frac: int) -> Optional[Frag]:
        """Decimal texts of k / 10**frac for the integers lo <= k <= hi (0 <= lo): an integer
        part without leading zeros, then (frac > 0) '.' and exactly `frac` digits."""
        if hi < lo:
            return None
        alts = []
        for total in range(max(len(str(lo)), frac + 1), max(len(str(hi)), frac + 1) + 1):
            # numbers whose text has `total` digits: a leading zero only as the lone integer digit
            first = 0 if total == frac + 1 else 10 ** (total - 1)
            a, z = max(lo, first), min(hi, 10 ** total - 1)
            if a <= z:
                alts.append(self.digits_between(str(a).zfill(total), str(z).zfill(total),
                                                total - frac if frac else None))
        if not alts:
            return None
        return alts[0] if len(alts) == 1 else self.alt(*alts)

    # ---- code-point sets inside a JSON string ------------------------------------
    def charset(self, ranges: Sequence[Tuple[int, int]]) -> Optional[Frag]:
        """One character out of a set of code points, spelled as it must appear inside a JSON
        string: the quote and the backslash escaped, control characters as their short escape
        (b f n r t) or the four-hex-digit form, the rest as UTF-8.  Surrogates are dropped."""
        b = self
        parts: List[Frag] = []
        plain = 0
        short = {0x08: b"\\b", 0x0C: b"\\f", 0x0A: b"\\n", 0x0D: b"\\r", 0x09: b"\\t"}
        for lo, hi in _normalise_ranges(ranges):
            for c in range(lo, min(hi, 0x7F) + 1):
                if c == 0x22:
                    parts.append(b.lit(b'\\"'))
                elif c == 0x5C:
                    parts.append(b.lit(b"\\\\"))
                elif c < 0x20:
                    parts.append(b.lit(short.get(c, b"\\u%04x" % c)))
                else:
                    plain |= 1 << c
            if hi >= 0x80:
                for seq in _utf8_sequences(max(lo, 0x80), hi):
                    parts.append(b.seq(*[b.bset(_mask((x, y))) for x, y in seq]))
        if plain:
            parts.append(b.bset(plain))
        if not parts:
            return None
        return parts[0] if len(parts) == 1 else b.alt(*parts)

    def embed(self, dfa: "ByteDFA") -> Frag:
        """A finished DFA as a fragment of this NFA (used for intersections)."""
        base = [self.n.new() for _ in range(dfa.n_states)]
        end = self.n.new()
        for q in range(dfa.n_states):
            by_target: Dict[int, int] = {}
            for byte in np.nonzero(dfa.trans[q] >= 0)[0]:
                t = int(dfa.trans[q, byte])
                by_target[t] = by_target.get(t, 0) | (1 << int(byte))
            for t, m in by_target.items():
                self.n.tr[base[q]].append((m, base[t]))
            if dfa.accept[q]:
                self.n.eps[base[q]].append(end)
        return (base[dfa.start], end)


_MAX_CP = 0x10FFFF


def _normalise_ranges(ranges: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """sorted, merged, without the surrogate block"""
    out: List[Tuple[int, int]] = []
    for lo, hi in sorted((max(0, a), min(_MAX_CP, z)) for a, z in ranges):
        if lo > hi:
            continue
        if out and lo <= out[-1][1] + 1:
            out[-1] = (out[-1][0], max(out[-1][1], hi))
        else:
            out.append((lo, hi))
    cut: List[Tuple[int, int]] = []
    for lo, hi in out:
        if hi < 0xD800 or lo > 0xDFFF:
            cut.append((lo, hi))
        else:
            if lo < 0xD800:
                cut.append((lo, 0xD7FF))
            if hi > 0xDFFF:
                cut.append((0xE000, hi))
    return cut


def _complement(ranges: Sequence[Tuple[int, int]], universe=(0, _MAX_CP)) -> List[Tuple[int, int]]:
    out, nxt = [], universe[0]
    for lo, hi in _normalise_ranges(ranges):
        if lo > nxt:
            out.append((nxt, lo - 1))
        nxt = max(nxt, hi + 1)
    if nxt <= universe[1]:
        out.append((nxt, universe[1]))
    return _normalise_ranges(out)


def _utf8_sequences(lo: int, hi: int) -> List[List[Tuple[int, int]]]:
    """Code points lo..hi (>= 0x80, no surrogates inside) as sequences of byte ranges."""
    out: List[List[Tuple[int, int]]] = []
    for n, (a, z) in ((2, (0x80, 0x7FF)), (3, (0x800, 0xFFFF)), (4, (0x10000, _MAX_CP))):
        x, y = max(lo, a), min(hi, z)
        if x <= y:
            _utf8_split(x, y, n, out)
    return out


def _utf8_split(lo: int, hi: int, n: int, out: List[List[Tuple[int, int]]]) -> None:
    for i in range(1, n):
        m = (1 << (6 * i)) - 1
        if (lo & ~m) != (hi & ~m):
            if lo & m:
                _utf8_split(lo, lo | m, n, out)
                _utf8_split((lo | m) + 1, hi, n, out)
                return
            if (hi & m) != m:
                _utf8_split(lo, (hi & ~m) - 1, n, out)
                _utf8_split(hi & ~m, hi, n, out)
                return
    a, z = chr(lo).encode("utf-8"), chr(hi).encode("utf-8")
    out.append([(a[k], z[k]) for k in range(n)])


# --------------------------------------------------------------------------- schema walk
class SchemaError(ValueError):
    pass


class _Compiler:
    def __init__(self, root: Dict[str, Any], limits: FsmLimits):
        self.root = root
        self.lim = limits
        self.b = _Builder()
        self._ref_stack: List[str] = []

    def resolve(self, ref: str) -> Dict[str, Any]:
        if not ref.startswith("#/"):
            raise SchemaError(f"unsupported $ref {ref!r}")
        node: Any = self.root
        for part in ref[2:].split("/"):
            node = node[part.replace("~1", "/").replace("~0", "~")]
        return node

    def lits(self, values) -> Frag:
        return self.b.literals([json.dumps(v, separators=(",", ":"), ensure_ascii=False)
                                .encode("utf-8") for v in values])

    # ---- numbers -----------------------------------------------------------
    @staticmethod
    def _bounds(sch):
        """-> (lo, lo_open, hi, hi_open) as Decimals / None"""
        def dec(v):
            return None if v is None else Decimal(str(v))
        lo, lo_open = dec(sch.get("minimum")), False
        hi, hi_open = dec(sch.get("maximum")), False
        xlo, xhi = dec(sch.get("exclusiveMinimum")), dec(sch.get("exclusiveMaximum"))
        if xlo is not None and (lo is None or xlo >= lo):
            lo, lo_open = xlo, True
        if xhi is not None and (hi is None or xhi <= hi):
            hi, hi_open = xhi, True
        return lo, lo_open, hi, hi_open

    def _signed(self, lo: Optional[int], hi: Optional[int], frac: int) -> Optional[Frag]:
        """texts of k / 10**frac for lo <= k <= hi (None = capped by max_int_digits)"""
        b = self.b
        cap = 10 ** (self.lim.max_int_digits + frac) - 1
        lo = -cap if lo is None else max(lo, -cap)
        hi = cap if hi is None else min(hi, cap)
        parts = []
        if hi >= 0:
            parts.append(b.scaled_range(max(lo, 0), hi, frac))
        if lo < 0:
            neg = b.scaled_range(max(-hi, 1), -lo, frac)     # "-0" is not produced
            if neg is not None:
                parts.append(b.seq(b.lit(b"-"), neg))
        parts = [f for f in parts if f is not None]
        if not parts:
            return None
        return parts[0] if len(parts) == 1 else b.alt(*parts)

    def _multiples(self, sch, lo, lo_open, hi, hi_open, integral: bool) -> Frag:
        step = Decimal(str(sch["multipleOf"]))
        if step <= 0:
            raise SchemaError("multipleOf must be positive")
        if lo is None or hi is None:
            raise SchemaError("multipleOf needs both a minimum and a maximum here")
        k0 = (lo / step).to_integral_value(rounding="ROUND_CEILING")
        k1 = (hi / step).to_integral_value(rounding="ROUND_FLOOR")
        if lo_open and k0 * step == lo:
            k0 += 1
        if hi_open and k1 * step == hi:
            k1 -= 1
        if k1 - k0 + 1 > self.lim.small_int_range:
            raise SchemaError("multipleOf over more than %d values is not supported"
                              % self.lim.small_int_range)
        words = []
        for k in range(int(k0), int(k1) + 1):
            v = k * step
            if integral and v != v.to_integral_value():
                continue
            text = format(v.normalize(), "f")
            words.append(("0" if text in ("-0", "") else text).encode())
        if not words:
            raise SchemaError("numeric range is empty")
        return self.b.literals(words)

    def integer(self, sch) -> Frag:
        lo, lo_open, hi, hi_open = self._bounds(sch)
        if sch.get("multipleOf") is not None:
            return self._multiples(sch, lo, lo_open, hi, hi_open, integral=True)
        ilo = ihi = None
        if lo is not None:
            ilo = int(lo.to_integral_value(rounding="ROUND_CEILING"))
            if lo_open and ilo == lo:
                ilo += 1
        if hi is not None:
            ihi = int(hi.to_integral_value(rounding="ROUND_FLOOR"))
            if hi_open and ihi == hi:
                ihi -= 1
        b = self.b
        if ilo is not None and ihi is not None:
            if ihi < ilo:
                raise SchemaError("integer range is empty")
            if ihi - ilo < self.lim.small_int_range:
                return b.literals([str(v).encode() for v in range(ilo, ihi + 1)])
        if ilo is None and ihi is None:
            digits = _mask((0x30, 0x39))
            body = b.alt(b.lit(b"0"), b.seq(b.bset(_mask((0x31, 0x39))),
                                           b.rep(lambda: b.bset(digits), 0,
                                                 self.lim.max_int_digits - 1)))
            return b.seq(b.opt(b.lit(b"-")), body)
        f = self._signed(ilo, ihi, 0)
        if f is None:
            raise SchemaError("integer range is empty (within %d digits)" % self.lim.max_int_digits)
        return f

    def number(self, sch) -> Frag:
        b = self.b
        lo, lo_open, hi, hi_open = self._bounds(sch)
        if sch.get("multipleOf") is not None:
            return self._multiples(sch, lo, lo_open, hi, hi_open, integral=False)
        if hi is None and (lo is None or (lo == 0 and not lo_open)):
            digits = _mask((0x30, 0x39))
            whole = b.alt(b.lit(b"0"), b.seq(b.bset(_mask((0x31, 0x39))),
                                            b.rep(lambda: b.bset(digits), 0,
                                                  self.lim.max_int_digits - 1)))
            frac = b.opt(b.seq(b.lit(b"."), b.rep(lambda: b.bset(digits), 1,
                                                  self.lim.max_frac_digits)))
            return b.seq(None if lo is not None else b.opt(b.lit(b"-")), whole, frac)
        # bounded: for every number of fraction digits f the admissible values are the
        # integers k with lo <= k / 10**f <= hi — an exact digit automaton per f
        parts = []
        for f in range(0, self.lim.max_frac_digits + 1):
            scale = Decimal(10) ** f
            klo = khi = None
            if lo is not None:
                klo = int((lo * scale).to_integral_value(rounding="ROUND_CEILING"))
                if lo_open and klo == lo * scale:
                    klo += 1
            if hi is not None:
                khi = int((hi * scale).to_integral_value(rounding="ROUND_FLOOR"))
                if hi_open and khi == hi * scale:
                    khi -= 1
            if klo is not None and khi is not None and khi < klo:
                continue
            frag = self._signed(klo, khi, f)
            if frag is not None:
                parts.append(frag)
        if not parts:
            raise SchemaError("numeric range is empty (within %d fraction digits)"
                              % self.lim.max_frac_digits)
        return parts[0] if len(parts) == 1 else b.alt(*parts)

    # ---- strings -------------------------------------------------------------
    _FORMATS = {
        # sound subsets of the formats pydantic validates (days stop at 28: valid in every month)
        "date": r"^[12]\d{3}-(0[1-9]|1[0-2])-(0[1-9]|1\d|2[0-8])$",
        "time": r"^([01]\d|2[0-3]):[0-5]\d:[0-5]\d$",
        "date-time": r"^[12]\d{3}-(0[1-9]|1[0-2])-(0[1-9]|1\d|2[0-8])T([01]\d|2[0-3]):[0-5]\d:[0-5]\dZ$",
        "uuid": r"^[0-9a-f]{8}-[0-9a-f]{4}-[1-5][0-9a-f]{3}-[89ab][0-9a-f]{3}-[0-9a-f]{12}$",
        "email": r"^[a-z0-9]{1,12}@[a-z0-9]{1,12}\.(com|org|net)$",
        "ipv4": r"^(25[0-5]|2[0-4]\d|1\d\d|[1-9]?\d)(\.(25[0-5]|2[0-4]\d|1\d\d|[1-9]?\d)){3}$",
        "uri": r"^https://[a-z0-9]{1,12}\.(com|org|net)(/[a-z0-9]{0,12})?$",
        "duration": r"^PT(\d{1,2}H)?(\d{1,2}M)?\d{1,2}S$",
    }
    # annotations: nothing to enforce
    _PLAIN_FORMATS = {"password", "binary", "byte", "regex", "path", "file-path",
                      "directory-path"}

    def string(self, sch) -> Frag:
        b = self.b
        lo = int(sch.get("minLength", 0))
        has_hi = sch.get("maxLength") is not None
        hi = int(sch["maxLength"]) if has_hi else max(lo, self.lim.max_string_chars)
        if hi < lo:
            raise SchemaError("string length range is empty")
        pattern = sch.get("pattern")
        fmt = sch.get("format")
        if fmt is not None and fmt not in self._PLAIN_FORMATS:
            if fmt not in self._FORMATS:
                raise SchemaError(f"unsupported string format {fmt!r}")
            if pattern is not None:
                raise SchemaError("format together with pattern is not supported")
            pattern = self._FORMATS[fmt]
        if pattern is None:
            return b.json_string(lo, hi)
        q = _mask(0x22)
        rb = _Builder()
        rx = _Regex(rb, self.lim)
        rf = rx.compile(pattern)
        rf = rb.seq(rb.bset(q), rf, rb.bset(q))
        pat = _determinise(rb.n, rf[0], rf[1])
        if lo == 0 and not has_hi and not rx.unbounded:
            return b.embed(pat)
        # length bounds, or loops that need the cap to terminate: product with the automaton
        # of "a JSON string of lo..hi characters" (loops stay loops, so this stays small)
        lb = _Builder()
        lf = lb.json_string(lo, hi)
        return b.embed(_intersect(pat, _determinise(lb.n, lf[0], lf[1])))

    # ---- arrays --------------------------------------------------------------
    def array(self, sch) -> Frag:
        b = self.b
        items = sch.get("items", {})
        prefix = sch.get("prefixItems")
        lo = int(sch.get("minItems", 0))
        has_hi = sch.get("maxItems") is not None
        if prefix is not None:
            return self._tuple(sch, list(prefix), items, lo,
                               int(sch["maxItems"]) if has_hi else None)
        hi = int(sch["maxItems"]) if has_hi else max(lo, self.lim.max_array_items)
        if hi < lo:
            raise SchemaError("array length range is empty")
        if sch.get("uniqueItems"):
            return self._unique(items, lo, hi)
        if hi == 0:
            return b.lit(b"[]")
        if items is False:
            if lo > 0:
                raise SchemaError("array admits no items but minItems > 0")
            return b.lit(b"[]")
        try:
            first = self.node(items)
        except SchemaError:
            if lo > 0:
                raise
            return b.lit(b"[]")      # items cannot be expressed (recursion floor): stay empty
        rest = b.rep(lambda: b.seq(b.lit(b","), self.node(items)), max(lo - 1, 0), hi - 1)
        inner = b.seq(first, rest)
        if lo == 0:
            inner = b.opt(inner)
        return b.seq(b.lit(b"["), inner, b.lit(b"]"))

    def _tuple(self, sch, prefix, items, lo: int, hi: Optional[int]) -> Frag:
        """prefixItems: positional schemas, then `items` for the rest (false = nothing more)."""
        b = self.b
        if sch.get("uniqueItems"):
            raise SchemaError("uniqueItems with prefixItems is not supported")
        n = len(prefix) if hi is None else min(len(prefix), hi)
        if lo > n and items is False:
            raise SchemaError("minItems exceeds the number of prefixItems")
        extra_lo = max(lo - n, 0)
        extra_hi = 0 if items is False else (
            (hi - n) if hi is not None else max(extra_lo, 0))   # no cap given: just the prefix
        alts = []
        # the positional part may stop early when minItems allows it
        for take in range(max(min(lo, n), 0), n + 1):
            if take < n and (extra_lo > 0):
                continue
            parts: List[Frag] = []
            for i in range(take):
                parts.append(b.lit(b"," if i else b""))
                parts.append(self.node(prefix[i]))
            if take == n and extra_hi > 0:
                more = b.rep(lambda: b.seq(b.lit(b"," if n else b""), self.node(items)),
                             extra_lo, extra_hi) if n else None
                if n == 0:
                    first = self.node(items)
                    rest = b.rep(lambda: b.seq(b.lit(b","), self.node(items)),
                                 max(extra_lo - 1, 0), extra_hi - 1)
                    more = b.seq(first, rest)
                    if extra_lo == 0:
                        more = b.opt(more)
                parts.append(more)
            parts = [p for p in parts if p is not None]
            alts.append(b.seq(b.lit(b"["), *parts, b.lit(b"]")) if parts else b.lit(b"[]"))
        return alts[0] if len(alts) == 1 else b.alt(*alts)

    def _unique(self, items, lo: int, hi: int) -> Frag:
        """uniqueItems needs memory a DFA does not have; small enumerations are spelled out."""
        values = None
        node = items
        if isinstance(node, dict) and "$ref" in node:
            node = self.resolve(node["$ref"])
        if isinstance(node, dict):
            if "enum" in node:
                values = list(node["enum"])
            elif "const" in node:
                values = [node["const"]]
            elif node.get("type") == "boolean":
                values = [True, False]
        if values is None or len(values) > 6:
            raise SchemaError("uniqueItems is only supported for arrays over at most 6 "
                              "enumerated values")
        seen, distinct = set(), []
        for v in values:
            key = json.dumps(v, sort_keys=True)
            if key not in seen:
                seen.add(key)
                distinct.append(v)
        arrays = [list(p) for k in range(lo, min(hi, len(distinct)) + 1)
                  for p in itertools.permutations(distinct, k)]
        if not arrays:
            raise SchemaError("uniqueItems: no array satisfies the length bounds")
        return self.lits(arrays)

    # ---- objects -------------------------------------------------------------
    def obj(self, sch) -> Frag:
        b = self.b
        props = sch.get("properties", {})
        addl = sch.get("additionalProperties")
        if not props:
            if isinstance(addl, dict) or addl is True and (
                    sch.get("minProperties") or sch.get("propertyNames")):
                return self._mapping(sch, addl if isinstance(addl, dict) else {})
            if int(sch.get("minProperties", 0)) > 0:
                raise SchemaError("minProperties > 0 on an object without properties")
            return b.lit(b"{}")
        if int(sch.get("minProperties", 0)) > len(props) or (
                sch.get("maxProperties") is not None and int(sch["maxProperties"]) < len(props)):
            raise SchemaError("min/maxProperties conflict with the listed properties")
        parts: List[Frag] = [b.lit(b"{")]
        for i, (name, sub) in enumerate(props.items()):
            key = json.dumps(name, ensure_ascii=False).encode("utf-8")
            parts.append(b.lit((b"," if i else b"") + key + b":"))
            parts.append(self.node(sub))
        parts.append(b.lit(b"}"))
        return b.seq(*parts)

    def _mapping(self, sch, value_schema) -> Frag:
        """Dict[str, T]: free keys (propertyNames applies), values of one schema.  Duplicate
        keys are not excluded (JSON allows them; parsers keep the last)."""
        b = self.b
        lo = int(sch.get("minProperties", 0))
        hi = int(sch["maxProperties"]) if sch.get("maxProperties") is not None \
            else max(lo, self.lim.max_array_items)
        if hi < lo:
            raise SchemaError("property count range is empty")
        names = dict(sch.get("propertyNames") or {})
        names.setdefault("type", "string")
        if names.get("type") != "string":
            raise SchemaError("propertyNames must describe strings")
        if "minLength" not in names and "pattern" not in names and "enum" not in names \
                and "format" not in names and "const" not in names:
            names["minLength"] = 1          # keep keys non-empty unless the schema says otherwise

        def entry():
            return b.seq(self.node(names), b.lit(b":"), self.node(value_schema))
        if hi == 0:
            return b.lit(b"{}")
        rest = b.rep(lambda: b.seq(b.lit(b","), entry()), max(lo - 1, 0), hi - 1)
        inner = b.seq(entry(), rest)
        if lo == 0:
            inner = b.opt(inner)
        return b.seq(b.lit(b"{"), inner, b.lit(b"}"))

    # ---- dispatch ------------------------------------------------------------
    _ANNOTATIONS = {"title", "description", "default", "examples", "example", "$schema", "$id",
                    "$comment", "$defs", "definitions", "deprecated", "readOnly", "writeOnly",
                    "discriminator", "contentEncoding", "contentMediaType", "nullable"}
    _KEYWORDS = {
        "object": {"properties", "required", "additionalProperties", "minProperties",
                   "maxProperties", "propertyNames"},
        "string": {"minLength", "maxLength", "pattern", "format"},
        "integer": {"minimum", "maximum", "exclusiveMinimum", "exclusiveMaximum", "multipleOf"},
        "number": {"minimum", "maximum", "exclusiveMinimum", "exclusiveMaximum", "multipleOf"},
        "array": {"items", "prefixItems", "minItems", "maxItems", "uniqueItems"},
        "boolean": set(), "null": set(),
    }

    def _check_keywords(self, sch, t: str) -> None:
        extra = set(sch) - self._ANNOTATIONS - self._KEYWORDS[t] - {"type"}
        if extra:
            raise SchemaError(f"unsupported keyword(s) {sorted(extra)} on a schema of type {t!r} "
                              "(they would constrain the output and cannot be ignored)")

    def any_value(self) -> Frag:
        # "any": keep it terminating — a string or a number or a literal
        return self.b.alt(self.string({}), self.number({}), self.lits([True, False, None]))

    def node(self, sch: Any) -> Frag:
        b = self.b
        if sch is True:
            return self.any_value()
        if not isinstance(sch, dict):
            raise SchemaError(f"unsupported schema node {sch!r}")
        if not (set(sch) - self._ANNOTATIONS):
            return self.any_value()
        if "$ref" in sch:
            ref = sch["$ref"]
            if set(sch) - self._ANNOTATIONS - {"$ref"}:
                raise SchemaError("keywords next to $ref are not supported")
            if self._ref_stack.count(ref) > self.lim.max_recursion:
                # a DFA has no stack: recursion is unrolled to a fixed depth; at the bottom the
                # enclosing array becomes [] / the enclosing union drops this alternative
                raise SchemaError(f"recursion through {ref} is deeper than "
                                  f"{self.lim.max_recursion} levels and nothing encloses it "
                                  "that could stop (an array with minItems 0, a union)")
            self._ref_stack.append(ref)
            try:
                return self.node(self.resolve(ref))
            finally:
                self._ref_stack.pop()
        if "const" in sch:
            return self.lits([sch["const"]])
        if "enum" in sch:
            if not sch["enum"]:
                raise SchemaError("enum is empty")
            return self.lits(sch["enum"])
        for k in ("anyOf", "oneOf"):
            if k in sch:
                if set(sch) - self._ANNOTATIONS - {k}:
                    raise SchemaError(f"keywords next to {k} are not supported")
                # an alternative this compiler cannot express is left out: the automaton
                # then accepts a subset of the union, which keeps every output valid
                alts, errors = [], []
                for x in sch[k]:
                    mark = (len(self.b.n.eps), len(self._ref_stack))
                    try:
                        alts.append(self.node(x))
                    except SchemaError as e:
                        errors.append(str(e))
                        del self._ref_stack[mark[1]:]
                if not alts:
                    raise SchemaError(f"no alternative of {k} is supported: " + "; ".join(errors))
                return alts[0] if len(alts) == 1 else b.alt(*alts)
        if "allOf" in sch:
            parts = list(sch["allOf"])
            rest = {k: v for k, v in sch.items() if k != "allOf"}
            if len(parts) == 1 and not (set(rest) - self._ANNOTATIONS):
                return self.node(parts[0])
            merged = dict(rest)
            for part in parts:
                if isinstance(part, dict) and "$ref" in part and len(parts) > 1:
                    part = self.resolve(part["$ref"])
                if not isinstance(part, dict):
                    raise SchemaError("allOf parts must be schema objects")
                for k, v in part.items():
                    if k in self._ANNOTATIONS:
                        continue
                    if k in merged and merged[k] != v:
                        raise SchemaError(f"allOf parts disagree on {k!r}: intersections of "
                                          "different constraints are not supported")
                    merged[k] = v
            return self.node(merged)
        for bad in ("not", "if", "then", "else", "contains", "patternProperties",
                    "dependentRequired", "dependentSchemas", "unevaluatedProperties",
                    "unevaluatedItems"):
            if bad in sch:
                raise SchemaError(f"unsupported keyword {bad!r}")
        t = sch.get("type")
        if isinstance(t, list):
            # {"type": ["string", "null"], "maxLength": 5}: every member type takes the keywords
            # that apply to it (JSON Schema: a keyword for another type is vacuous)
            typed = set().union(*self._KEYWORDS.values())
            alts = []
            for x in t:
                if x not in self._KEYWORDS:
                    raise SchemaError(f"unsupported type {x!r}")
                keep = {k: v for k, v in sch.items()
                        if k not in typed or k in self._KEYWORDS[x]}
                alts.append(self.node({**keep, "type": x}))
            return alts[0] if len(alts) == 1 else b.alt(*alts)
        if t is None:
            for name, keys in self._KEYWORDS.items():
                if keys & set(sch):
                    t = name if name != "integer" else "number"
                    break
        if t is None:
            raise SchemaError(f"unsupported schema node {sch!r}")
        if t not in self._KEYWORDS:
            raise SchemaError(f"unsupported type {t!r}")
        self._check_keywords(sch, t)
        if t == "object":
            return self.obj(sch)
        if t == "string":
            return self.string(sch)
        if t == "integer":
            return self.integer(sch)
        if t == "number":
            return self.number(sch)
        if t == "boolean":
            return self.lits([True, False])
        if t == "null":
            return self.lits([None])
        return self.array(sch)


# --------------------------------------------------------------------------- patterns
class _Regex:
    """`pattern` -> fragment over JSON-escaped UTF-8.  The pattern is parsed by Python's own
    regex parser; supported: literals, classes (ranges, \\d \\w \\s and their negations), '.',
    groups, alternation, greedy/lazy repeats (unbounded ones capped by FsmLimits), ^ and $ at
    the ends.  JSON Schema patterns are unanchored: an open end is padded with "any characters".
    \\d, \\w, \\s mean the ASCII sets and negated classes that mention them stay inside ASCII —
    subsets of both the ECMA-262 and the Unicode-aware reading, so validators of either kind
    accept what this automaton produces."""

    _DIGIT = [(0x30, 0x39)]
    _WORD = [(0x30, 0x39), (0x41, 0x5A), (0x5F, 0x5F), (0x61, 0x7A)]
    _SPACE = [(0x09, 0x0D), (0x20, 0x20)]
    _ASCII = (0x00, 0x7F)

    def __init__(self, b: _Builder, lim: FsmLimits):
        self.b, self.lim = b, lim
        self.unbounded = False      # a loop was used: the caller must bound the total length

    def compile(self, pattern: str) -> Frag:
        import re._constants as K
        import re._parser as P
        try:
            tree = P.parse(pattern)
        except Exception as e:
            raise SchemaError(f"pattern {pattern!r} does not parse: {e}")
        if tree.state.flags & ~(tree.state.flags & 32):       # anything but re.UNICODE
            raise SchemaError("pattern flags are not supported")
        self.K = K
        items = list(tree)
        open_left = not (items and items[0][0] is K.AT and
                         items[0][1] in (K.AT_BEGINNING, K.AT_BEGINNING_STRING))
        if not open_left:
            items = items[1:]
        open_right = not (items and items[-1][0] is K.AT and
                          items[-1][1] in (K.AT_END, K.AT_END_STRING))
        if not open_right:
            items = items[:-1]
        b = self.b
        if open_left or open_right:
            self.unbounded = True
        pad = lambda: b.star(b.json_char)   # noqa: E731
        parts = [pad() if open_left else None, self.seq(items), pad() if open_right else None]
        parts = [p for p in parts if p is not None]
        return b.seq(*parts) if parts else b.empty()

    def seq(self, items) -> Optional[Frag]:
        frags = [self.one(op, arg) for op, arg in items]
        frags = [f for f in frags if f is not None]
        return self.b.seq(*frags) if frags else None

    def category(self, cat):
        K = self.K
        table = {K.CATEGORY_DIGIT: (self._DIGIT, False), K.CATEGORY_NOT_DIGIT: (self._DIGIT, True),
                 K.CATEGORY_WORD: (self._WORD, False), K.CATEGORY_NOT_WORD: (self._WORD, True),
                 K.CATEGORY_SPACE: (self._SPACE, False), K.CATEGORY_NOT_SPACE: (self._SPACE, True)}
        if cat not in table:
            raise SchemaError(f"unsupported character category {cat}")
        rs, neg = table[cat]
        return (_complement(rs, self._ASCII) if neg else list(rs)), True

    def char_class(self, members) -> List[Tuple[int, int]]:
        K = self.K
        negate, ranges, ascii_only = False, [], False
        for op, arg in members:
            if op is K.NEGATE:
                negate = True
            elif op is K.LITERAL:
                ranges.append((arg, arg))
            elif op is K.RANGE:
                ranges.append((arg[0], arg[1]))
            elif op is K.CATEGORY:
                rs, a = self.category(arg)
                ranges += rs
                ascii_only = ascii_only or a
            else:
                raise SchemaError(f"unsupported item in a character class: {op}")
        if negate:
            return _complement(ranges, self._ASCII if ascii_only else (0, _MAX_CP))
        return ranges

    def chars(self, ranges) -> Frag:
        f = self.b.charset(ranges)
        if f is None:
            raise SchemaError("pattern contains an empty character class")
        return f

    def one(self, op, arg) -> Optional[Frag]:
        K, b = self.K, self.b
        if op is K.LITERAL:
            return self.chars([(arg, arg)])
        if op is K.NOT_LITERAL:
            return self.chars(_complement([(arg, arg)]))
        if op is K.ANY:
            return self.chars(_complement([(0x0A, 0x0A)]))
        if op is K.IN:
            return self.chars(self.char_class(arg))
        if op is K.CATEGORY:
            return self.chars(self.category(arg)[0])
        if op is K.SUBPATTERN:
            _group, add_flags, del_flags, sub = arg
            if add_flags or del_flags:
                raise SchemaError("inline pattern flags are not supported")
            return self.seq(list(sub))
        if op is K.BRANCH:
            alts = [self.seq(list(x)) or b.empty() for x in arg[1]]
            return b.alt(*alts)
        if op in (K.MAX_REPEAT, K.MIN_REPEAT):
            lo, hi, sub = arg
            sub = list(sub)
            if self.seq(sub) is None:
                return None
            if lo > 4096:
                raise SchemaError("pattern repeat count is too large")
            if hi is K.MAXREPEAT or hi - lo > 256:
                # a loop, not copies; the string's length cap keeps generation finite
                self.unbounded = True
                head = b.rep(lambda: self.seq(sub), lo, lo) if lo else None
                return b.seq(head, b.star(lambda: self.seq(sub)))
            return b.rep(lambda: self.seq(sub), lo, hi)
        if op is K.AT:
            raise SchemaError("anchors and boundaries inside a pattern are not supported")
        raise SchemaError(f"unsupported pattern construct {op}")


def _intersect(a: "ByteDFA", c: "ByteDFA") -> "ByteDFA":
    """product automaton (reachable part), trimmed like _determinise's output"""
    ids: Dict[Tuple[int, int], int] = {(a.start, c.start): 0}
    order = [(a.start, c.start)]
    rows: List[np.ndarray] = []
    i = 0
    while i < len(order):
        p, q = order[i]
        i += 1
        row = np.full(256, -1, dtype=np.int32)
        ta, tc = a.trans[p], c.trans[q]
        for byte in np.nonzero((ta >= 0) & (tc >= 0))[0]:
            key = (int(ta[byte]), int(tc[byte]))
            if key not in ids:
                if len(order) >= _MAX_DFA_STATES:
                    raise SchemaError("the schema's automaton exceeds %d states" % _MAX_DFA_STATES)
                ids[key] = len(order)
                order.append(key)
            row[byte] = ids[key]
        rows.append(row)
    trans = np.stack(rows)
    accept = np.asarray([a.accept[p] and c.accept[q] for p, q in order], dtype=np.uint8)
    live = accept.astype(bool).copy()
    while True:
        nxt = live | ((trans >= 0) & live[np.clip(trans, 0, None)]).any(axis=1)
        if (nxt == live).all():
            break
        live = nxt
    if not live[0]:
        raise SchemaError("pattern and length bounds admit no string")
    trans = np.where((trans >= 0) & live[np.clip(trans, 0, None)], trans, -1).astype(np.int32)
    final = (accept.astype(bool) & (trans < 0).all(axis=1)).astype(np.uint8)
    return ByteDFA(np.ascontiguousarray(trans), accept, final, 0)


# --------------------------------------------------------------------------- NFA -> DFA
# The mask-build kernel runs one grid row per state (grid.y <= 65535), and every state costs
# vocab/8 bytes of mask in HBM (19 KB at a 152 k vocabulary); larger automata are refused.
_MAX_DFA_STATES = 65_535


def _determinise(nfa: _NFA, start: int, end: int) -> ByteDFA:
    n = len(nfa.eps)
    # byte equivalence classes: bytes that no transition mask distinguishes
    masks = sorted({m for trs in nfa.tr for m, _ in trs})
    cls_of = np.zeros(256, dtype=np.int32)
    keys: Dict[Tuple[bool, ...], int] = {}
    for byte in range(256):
        k = tuple(bool((m >> byte) & 1) for m in masks)
        if k not in keys:
            keys[k] = len(keys)
        cls_of[byte] = keys[k]
    n_cls = len(keys)
    rep_byte = [int(np.argmax(cls_of == c)) for c in range(n_cls)]

    def closure(states) -> frozenset:
        out, stack = set(states), list(states)
        while stack:
            s = stack.pop()
            for t in nfa.eps[s]:
                if t not in out:
                    out.add(t)
                    stack.append(t)
        return frozenset(out)

    # per NFA state: class -> targets
    step: List[Dict[int, List[int]]] = []
    for s in range(n):
        d: Dict[int, List[int]] = {}
        for m, t in nfa.tr[s]:
            for c in range(n_cls):
                if (m >> rep_byte[c]) & 1:
                    d.setdefault(c, []).append(t)
        step.append(d)

    start_set = closure([start])
    ids: Dict[frozenset, int] = {start_set: 0}
    order = [start_set]
    rows: List[List[int]] = []
    i = 0
    while i < len(order):
        cur = order[i]
        i += 1
        row = [-1] * n_cls
        moved: Dict[int, set] = {}
        for s in cur:
            for c, ts in step[s].items():
                moved.setdefault(c, set()).update(ts)
        for c, ts in moved.items():
            nxt = closure(ts)
            if nxt not in ids:
                if len(order) >= _MAX_DFA_STATES:
                    raise SchemaError("the schema's automaton exceeds %d states (counted "
                                      "repeats nested in alternations blow up when determinised)"
                                      % _MAX_DFA_STATES)
                ids[nxt] = len(order)
                order.append(nxt)
            row[c] = ids[nxt]
        rows.append(row)
    cls_trans = np.asarray(rows, dtype=np.int32).reshape(len(order), n_cls)
    accept = np.asarray([end in st for st in order], dtype=np.uint8)

    # trim: drop states that cannot reach an accepting state
    live = accept.astype(bool).copy()
    changed = True
    while changed:
        nxt_live = live.copy()
        for c in range(n_cls):
            t = cls_trans[:, c]
            nxt_live |= (t >= 0) & live[np.clip(t, 0, None)]
        changed = bool((nxt_live != live).any())
        live = nxt_live
    cls_trans = np.where((cls_trans >= 0) & live[np.clip(cls_trans, 0, None)], cls_trans, -1)
    if not live[0]:
        raise SchemaError("schema accepts nothing")
    trans = cls_trans[:, cls_of]            # expand classes back to bytes
    final = (accept.astype(bool) & (trans < 0).all(axis=1)).astype(np.uint8)
    return ByteDFA(np.ascontiguousarray(trans, dtype=np.int32), accept, final, 0)


def compile_schema(schema: Dict[str, Any], limits: Optional[FsmLimits] = None) -> ByteDFA:
    """JSON schema (dict) -> ByteDFA.  Raises SchemaError (a ValueError, like the
    reference's own argument errors) for constructs outside the supported subset."""
    if not isinstance(schema, dict):
        raise SchemaError("schema must be a dict (use normalize_output_schema first)")
    comp = _Compiler(schema, limits or FsmLimits())
    frag = comp.node(schema)
    return _determinise(comp.b.n, frag[0], frag[1])


THINK_CLOSE = b"</think>\n\n"


def compile_thinking(schema: Optional[Dict[str, Any]], limits: Optional[FsmLimits] = None,
                     think_chars: int = 128) -> ByteDFA:
    """Automaton of a thinking model's turn: up to `think_chars` characters of free reasoning
    (any text without '<'), the closing `</think>` line, then the content — an instance of
    `schema`, or free text when there is none.  The reference reports such outputs as
    `{"content": ..., "reasoning_content": ...}` (sutro/sdk.py:1155-1164); the cap plays the
    role FsmLimits plays for strings: with it every row reaches its content."""
    lim = limits or FsmLimits()
    comp = _Compiler(schema if isinstance(schema, dict) else {}, lim)
    b = comp.b
    plain = _mask(0x09, 0x0A, (0x20, 0x3B), (0x3D, 0x7E))

    def text_char(allow_lt: bool):
        parts = [b.bset(plain | (_mask(0x3C) if allow_lt else 0))]
        for lo, hi in ((0x80, 0xD7FF), (0xE000, _MAX_CP)):
            for seq in _utf8_sequences(lo, hi):
                parts.append(b.seq(*[b.bset(_mask((x, y))) for x, y in seq]))
        return b.alt(*parts)
    reasoning = b.rep(lambda: text_char(False), 0, int(think_chars))
    close = b.lit(THINK_CLOSE)
    if schema is not None:
        content = comp.node(schema)
    else:
        content = b.star(lambda: text_char(True))
    frag = b.seq(reasoning, close, content)
    return _determinise(b.n, frag[0], frag[1])


def split_thinking(text: str):
    """-> (reasoning_content, content) of a thinking turn produced under compile_thinking."""
    head, sep, tail = text.partition("</think>")
    if not sep:
        return text.strip(), ""
    return head.strip(), tail.lstrip("\n")
