"""sb200_schema_compile (csrc/schema_compile.cu: JSON Schema text -> byte DFA behind the C-ABI)
against the Python compiler (sutro_b200/schema_fsm.py) — the two automata must accept EXACTLY
the same language (checked by walking the product automaton: both are deterministic and
trimmed, so equivalence is a bisimulation from the start states), on the schemas the
reference's templates and tests use (sutro/templates/classification.py:83-89,
evals.py:42-52, :107-121, tests/test_sdk.py:427-435, README.md:45-46) and on Pydantic output.
Runs without a GPU: the compiler is host code inside libsutro_b200.so."""
import ctypes as C
import json
from typing import List, Literal, Optional

import numpy as np
import pytest
from pydantic import BaseModel, Field

from sutro_b200 import _lib as L
from sutro_b200 import engine  # noqa: F401  (registers the engine-level symbols)
from sutro_b200.schema_fsm import ByteDFA, FsmLimits, SchemaError, compile_schema


def native(schema, limits: FsmLimits = None) -> ByteDFA:
    from sutro_b200.engine import FsmLimitsC, native_compile_schema
    return native_compile_schema(schema, limits)


def equivalent(a: ByteDFA, b: ByteDFA):
    """None when the languages are equal, else a witness string accepted by exactly one."""
    seen = {(a.start, b.start): b""}
    stack = [(a.start, b.start)]
    while stack:
        p, q = stack.pop()
        w = seen[(p, q)]
        if bool(a.accept[p]) != bool(b.accept[q]):
            return w
        ta, tb = a.trans[p], b.trans[q]
        if ((ta >= 0) != (tb >= 0)).any():
            byte = int(np.nonzero((ta >= 0) != (tb >= 0))[0][0])
            return w + bytes([byte])
        for byte in np.nonzero(ta >= 0)[0]:
            nxt = (int(ta[byte]), int(tb[byte]))
            if nxt not in seen:
                seen[nxt] = w + bytes([int(byte)]) if len(w) < 64 else w
                stack.append(nxt)
    return None


class Sentiment(BaseModel):
    sentiment: Literal["positive", "neutral", "negative"]


class FreeSentiment(BaseModel):
    sentiment: str


class Item(BaseModel):
    name: str = Field(max_length=12)
    quantity: int = Field(ge=0, le=1000)
    kind: Literal["a", "b", "c"]
    price: Optional[float] = None


class Order(BaseModel):
    customer: str = Field(max_length=10)
    items: List[Item] = Field(max_length=3)
    paid: bool


class Node(BaseModel):
    value: int = Field(ge=-5, le=5)
    children: List["Node"] = Field(default_factory=list, max_length=2)


class Event(BaseModel):
    when: __import__("datetime").datetime
    day: __import__("datetime").date
    ident: __import__("uuid").UUID
    pair: __import__("typing").Tuple[int, bool]
    tags: __import__("typing").Set[Literal["a", "b"]]
    counts: __import__("typing").Dict[str, bool]


SMALL = FsmLimits(max_string_chars=6, max_array_items=2, max_int_digits=4, max_frac_digits=2)
CASES = [
    (Sentiment.model_json_schema(), None),
    (FreeSentiment.model_json_schema(), FsmLimits(max_string_chars=10)),
    (Order.model_json_schema(), FsmLimits(max_string_chars=12, max_array_items=3)),
    (Node.model_json_schema(), SMALL),
    ({"type": "object", "properties": {"scratchpad": {"type": "string", "maxLength": 8},
                                       "classification": {"type": "string", "enum": ["x", "y z", "é\"q"]}},
      "required": ["scratchpad", "classification"]}, None),
    ({"type": "object", "properties": {"score": {"type": "integer", "minimum": 0, "maximum": 10}},
      "required": ["score"]}, None),
    ({"type": "object", "properties": {"ranking": {"type": "array", "items": {"type": "string", "maxLength": 3},
                                                   "minItems": 1, "maxItems": 3}}}, None),
    ({"type": "integer", "minimum": -1234, "maximum": 99999}, None),
    ({"type": "integer", "exclusiveMinimum": 5}, SMALL),
    ({"type": "integer"}, SMALL),
    ({"type": "number", "minimum": -2.5, "exclusiveMaximum": 10}, SMALL),
    ({"type": "number", "minimum": 0}, SMALL),
    ({"type": "number"}, SMALL),
    ({"type": ["string", "null"], "maxLength": 3}, None),
    ({"anyOf": [{"type": "boolean"}, {"type": "null"}, {"const": {"a": [1, 2.5, "x"], "b": None}}]}, None),
    ({"enum": [1, 2.0, "three", True, None, [1, "a"], {"k": "v"}, 1e22, -0.5]}, None),
    ({"allOf": [{"type": "string"}, {"maxLength": 2}], "minLength": 1}, None),
    ({"type": "array", "items": {"type": "boolean"}}, SMALL),
    ({"type": "array", "items": False}, None),
    ({"type": "object"}, None),
    ({}, SMALL),
    ({"type": "object", "properties": {"kéy \"q\"": {"type": "null"}}}, None),
    # what Pydantic emits for datetime / date / time / UUID, Tuple, Set and Dict fields
    (Event.model_json_schema(), SMALL),
    ({"type": "string", "format": "date"}, None),
    ({"type": "string", "format": "time"}, None),
    ({"type": "string", "format": "date-time"}, None),
    ({"type": "string", "format": "uuid"}, None),
    ({"type": "string", "format": "email"}, None),
    ({"type": "string", "format": "uri"}, None),
    ({"type": "string", "format": "ipv4"}, None),
    ({"type": "string", "format": "duration"}, None),
    ({"type": "array", "prefixItems": [{"type": "integer", "minimum": 0, "maximum": 9}, {"type": "boolean"}],
      "items": False}, None),
    ({"type": "array", "prefixItems": [{"type": "boolean"}, {"type": "null"}], "minItems": 1,
      "items": {"type": "integer", "minimum": 0, "maximum": 3}, "maxItems": 4}, None),
    ({"type": "array", "prefixItems": [{"enum": ["a", "b"]}], "minItems": 0}, None),
    ({"type": "array", "prefixItems": [], "items": {"type": "boolean"}, "maxItems": 2}, None),
    ({"type": "array", "items": {"enum": [1, "two", None]}, "uniqueItems": True}, None),
    ({"type": "array", "items": {"type": "boolean"}, "uniqueItems": True, "minItems": 1}, None),
    ({"type": "array", "items": {"const": {"k": [1]}}, "uniqueItems": True}, None),
    ({"type": "object", "additionalProperties": {"type": "integer", "minimum": 0, "maximum": 5}}, SMALL),
    ({"type": "object", "additionalProperties": {"type": "boolean"}, "minProperties": 1, "maxProperties": 2,
      "propertyNames": {"enum": ["x", "y"]}}, None),
    ({"type": "object", "additionalProperties": True, "minProperties": 1,
      "propertyNames": {"maxLength": 2}}, SMALL),
    ({"type": "integer", "minimum": 0, "maximum": 10, "multipleOf": 2}, None),
    ({"type": "integer", "minimum": -7, "exclusiveMaximum": 9, "multipleOf": 3}, None),
    ({"type": "integer", "minimum": -3, "maximum": 3, "multipleOf": 1.5}, None),
    ({"type": "number", "minimum": -1, "maximum": 1, "multipleOf": 0.25}, None),
    ({"type": "number", "exclusiveMinimum": 0, "maximum": 0.3, "multipleOf": 0.05}, None),
    ({"type": "number", "minimum": 10, "maximum": 100, "multipleOf": 12.5}, None),
]


@pytest.mark.parametrize("schema,limits", CASES)
def test_native_schema_compiler_accepts_the_same_language(schema, limits):
    want = compile_schema(schema, limits)
    got = native(schema, limits)
    witness = equivalent(want, got)
    assert witness is None, (witness, want.matches(witness), got.matches(witness))
    assert got.longest_path() == want.longest_path()
    # the jump-forward plan is a function of the language: same forced prefix, same tails
    assert got.forced_plan()[0] == want.forced_plan()[0]
    assert sorted(got.forced_plan()[2].values()) == sorted(want.forced_plan()[2].values())


@pytest.mark.parametrize("schema", [
    {"type": "string", "pattern": "^[a-z]+$"}, {"type": "string", "format": "hostname"},
    {"type": "string", "format": "date", "maxLength": 10},
    {"type": "integer", "minimum": 0, "multipleOf": 2}, {"type": "number", "minimum": 0, "maximum": 1, "multipleOf": 0},
    {"type": "array", "items": {"type": "integer"}, "uniqueItems": True},
    {"type": "array", "prefixItems": [{"type": "integer"}], "uniqueItems": True},
    {"type": "object", "additionalProperties": {"type": "integer"}, "propertyNames": {"pattern": "^k"}},
    {"type": "string", "bogusKeyword": 1}, {"not": {"type": "string"}}, {"$ref": "http://x/y"},
    {"type": "integer", "minimum": 5, "maximum": 1},
])
def test_native_compiler_refuses_what_it_does_not_implement(schema):
    """Nothing that would constrain the output is silently ignored (argument error, like the
    Python compiler's SchemaError / the reference's ValueError convention)."""
    with pytest.raises(ValueError):
        native(schema)
    text = json.dumps(schema).encode()
    h = C.c_void_p()
    assert L.lib().sb200_schema_compile(text, len(text), None, C.byref(h)) == -2
    assert b"output_schema" in L.lib().sb200_last_error()


def test_native_union_drops_only_the_alternative_it_cannot_express():
    """anyOf / oneOf: an alternative outside the native subset (here: a pattern) is left out —
    the automaton accepts a SUBSET of the union, so every output stays valid (same policy as
    the Python compiler for alternatives it cannot express)."""
    schema = {"oneOf": [{"type": "string", "pattern": "^a+$"},
                        {"type": "integer", "minimum": 1, "maximum": 3}]}
    got, want = native(schema), compile_schema(schema)
    assert got.matches(b"2") and not got.matches(b'"aa"') and want.matches(b'"aa"')
    assert equivalent(got, compile_schema({"type": "integer", "minimum": 1, "maximum": 3})) is None


def test_native_compiler_rejects_bad_json_and_reports_limits():
    h = C.c_void_p()
    for bad in (b"", b"{", b"[1,2]", b'{"type": "string"} x', b'{"a": tru}'):
        assert L.lib().sb200_schema_compile(bad, len(bad), None, C.byref(h)) != 0
    from sutro_b200.engine import FsmLimitsC
    lim = FsmLimitsC()
    L.lib().sb200_fsm_limits_default(C.byref(lim))
    d = FsmLimits()
    assert (lim.max_string_chars, lim.max_array_items, lim.max_int_digits, lim.max_frac_digits,
            lim.small_int_range, lim.max_recursion) == (
        d.max_string_chars, d.max_array_items, d.max_int_digits, d.max_frac_digits,
        d.small_int_range, d.max_recursion)
