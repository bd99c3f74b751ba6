"""output_schema -> byte-level DFA for constrained decoding (host side of K8).

The reference only normalises the schema (sutro/common.py:152-163: a Pydantic class
becomes `.model_json_schema()`, a dict passes through) and ships it as `json_schema`
in the request (sutro/sdk.py:199); enforcing it is the server's job.  Locally the
schema is compiled to a deterministic automaton over UTF-8 bytes that accepts exactly
the *compact* JSON serialisations (no optional white space, properties in declaration
order) of instances of the schema; the GPU turns it into per-state token masks
(csrc/sampler_fsm.cu).

Supported subset (what Pydantic emits for plain models): object/properties (all
listed properties are emitted, in order), string (minLength/maxLength, enum, const),
integer / number (minimum/maximum; exact when the range is small), boolean, null,
enum of JSON literals, array (items, minItems/maxItems), anyOf/oneOf, $ref/$defs
(non-recursive).  Unbounded strings / arrays / digits get explicit caps
(`FsmLimits`) so that every path through the automaton terminates — with random
weights a model never chooses to stop on its own.
"""
from __future__ import annotations

import json
from dataclasses import dataclass
from typing import Any, Callable, Dict, List, Optional, Tuple

import numpy as np


@dataclass(frozen=True)
class FsmLimits:
    max_string_chars: int = 64     # cap when a string has no maxLength
    max_array_items: int = 8       # cap when an array has no maxItems
    max_int_digits: int = 9        # digits of an unbounded integer
    max_frac_digits: int = 4
    small_int_range: int = 2048    # ranges up to this size are encoded exactly


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

    def integer(self, sch) -> Frag:
        lo, hi = sch.get("minimum"), sch.get("maximum")
        if sch.get("exclusiveMinimum") is not None:
            lo = int(sch["exclusiveMinimum"]) + 1
        if sch.get("exclusiveMaximum") is not None:
            hi = int(sch["exclusiveMaximum"]) - 1
        if lo is not None and hi is not None:
            lo, hi = int(lo), int(hi)
            if hi < lo:
                raise SchemaError("integer range is empty")
            if hi - lo < self.lim.small_int_range:
                return self.b.literals([str(v).encode() for v in range(lo, hi + 1)])
        b = self.b
        nd = self.lim.max_int_digits
        if hi is not None and hi >= 0:
            nd = min(nd, len(str(int(hi))))
        digits = _mask((0x30, 0x39))
        body = b.alt(b.lit(b"0"), b.seq(b.bset(_mask((0x31, 0x39))),
                                       b.rep(lambda: b.bset(digits), 0, nd - 1)))
        neg_ok = lo is None or lo < 0
        return b.seq(b.opt(b.lit(b"-")), body) if neg_ok else body

    def number(self, sch) -> Frag:
        b = self.b
        digits = _mask((0x30, 0x39))
        whole = b.alt(b.lit(b"0"), b.seq(b.bset(_mask((0x31, 0x39))),
                                        b.rep(lambda: b.bset(digits), 0,
                                              self.lim.max_int_digits - 1)))
        frac = b.opt(b.seq(b.lit(b"."), b.rep(lambda: b.bset(digits), 1, self.lim.max_frac_digits)))
        lo = sch.get("minimum", sch.get("exclusiveMinimum"))
        sign = None if (lo is not None and lo >= 0) else b.opt(b.lit(b"-"))
        return b.seq(sign, whole, frac)

    def string(self, sch) -> Frag:
        lo = int(sch.get("minLength", 0))
        hi = int(sch.get("maxLength", max(lo, self.lim.max_string_chars)))
        if hi < lo:
            raise SchemaError("string length range is empty")
        return self.b.json_string(lo, hi)

    def array(self, sch) -> Frag:
        b = self.b
        items = sch.get("items", {})
        lo = int(sch.get("minItems", 0))
        hi = int(sch.get("maxItems", max(lo, self.lim.max_array_items)))
        if hi == 0:
            return b.lit(b"[]")
        first = self.node(items)
        rest = b.rep(lambda: b.seq(b.lit(b","), self.node(items)), max(lo - 1, 0), hi - 1)
        inner = b.seq(first, rest)
        if lo == 0:
            inner = b.opt(inner)
        return b.seq(b.lit(b"["), inner, b.lit(b"]"))

    def obj(self, sch) -> Frag:
        b = self.b
        props = sch.get("properties", {})
        if not props:
            return b.lit(b"{}")
        parts: List[Frag] = [b.lit(b"{")]
        for i, (name, sub) in enumerate(props.items()):
            key = json.dumps(name, ensure_ascii=False).encode("utf-8")
            parts.append(b.lit((b"," if i else b"") + key + b":"))
            parts.append(self.node(sub))
        parts.append(b.lit(b"}"))
        return b.seq(*parts)

    def node(self, sch: Any) -> Frag:
        b = self.b
        if sch is True or sch == {}:
            # "any": keep it terminating — a string or a number or a literal
            return b.alt(self.string({}), self.number({}), self.lits([True, False, None]))
        if not isinstance(sch, dict):
            raise SchemaError(f"unsupported schema node {sch!r}")
        if "$ref" in sch:
            ref = sch["$ref"]
            if ref in self._ref_stack:
                raise SchemaError(f"recursive schema through {ref} is not supported")
            self._ref_stack.append(ref)
            try:
                return self.node(self.resolve(ref))
            finally:
                self._ref_stack.pop()
        if "const" in sch:
            return self.lits([sch["const"]])
        if "enum" in sch:
            return self.lits(sch["enum"])
        for k in ("anyOf", "oneOf"):
            if k in sch:
                return b.alt(*[self.node(s) for s in sch[k]])
        if "allOf" in sch and len(sch["allOf"]) == 1:
            return self.node(sch["allOf"][0])
        t = sch.get("type")
        if isinstance(t, list):
            return b.alt(*[self.node({**sch, "type": x}) for x in t])
        if t == "object" or (t is None and "properties" in sch):
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
        if t == "array":
            return self.array(sch)
        raise SchemaError(f"unsupported schema node {sch!r}")


# --------------------------------------------------------------------------- NFA -> DFA
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
