"""The wider JSON-Schema subset of the schema -> DFA compiler: formats, patterns, exact
numeric ranges, tuples, mappings, uniqueItems.

Soundness: random strings accepted by the automaton must validate under pydantic (the
validator the reference's users bring, sutro/common.py:152-163) — i.e. nothing that
constrains the output is silently ignored.  Completeness: hand-picked valid instances are
accepted.  Unsupported constraining keywords raise SchemaError (a ValueError)."""
import datetime
import enum
import json
import re
import uuid
from decimal import Decimal
from typing import Any, Dict, List, Literal, Optional, Tuple, Union

import numpy as np
import pytest
from pydantic import BaseModel, Field

from oracle.fsm_ref import random_accepted
from sutro_b200.schema_fsm import FsmLimits, SchemaError, compile_schema

LIM = FsmLimits(max_string_chars=8, max_array_items=3)


class Color(enum.Enum):
    red = "red"
    blue = "blue"


class Mapping(BaseModel):
    d: Dict[str, int]


class Pair(BaseModel):
    t: Tuple[int, str]
    u: Tuple[bool, Literal["x", "y"], float] = (True, "x", 0.5)


class Dates(BaseModel):
    day: datetime.date
    at: datetime.datetime
    clock: datetime.time
    id: uuid.UUID


class OpenInterval(BaseModel):
    x: int = Field(gt=0, lt=10)
    y: int = Field(ge=-3000, le=70000)
    z: int = Field(ge=5)
    w: int = Field(le=-17)


class Pattern(BaseModel):
    code: str = Field(pattern=r"^[a-z]{2,4}-\d+$")
    loose: str = Field(pattern=r"ab+c")
    both: str = Field(pattern=r"^\w+$", min_length=3, max_length=5)


class Floats(BaseModel):
    p: float = Field(ge=0.0, le=1.0)
    q: float = Field(gt=-2.5, lt=2.5)
    r: float = Field(ge=10.25)
    s: float = Field(lt=0)


class Multiples(BaseModel):
    m: int = Field(multiple_of=5, ge=0, le=50)
    n: float = Field(multiple_of=0.25, ge=0, le=2)
    c: Color


class Nested(BaseModel):
    s: str = Field(min_length=2, max_length=4)
    o: Optional[int] = None
    l: List[List[int]] = Field(max_length=2)
    u: Union[int, str, None]
    lit: Literal[1, "a", True, None]
    anything: Any = None


MODELS = [Mapping, Pair, Dates, OpenInterval, Pattern, Floats, Multiples, Nested]


@pytest.mark.parametrize("model", MODELS, ids=lambda m: m.__name__)
def test_every_accepted_string_validates_under_pydantic(model):
    dfa = compile_schema(model.model_json_schema(), LIM)
    rng = np.random.RandomState(1)
    for _ in range(250):
        s = random_accepted(dfa, rng)
        model.model_validate_json(s)          # raises on a schema violation


def accepts(schema, value, lim=LIM) -> bool:
    text = json.dumps(value, separators=(",", ":"), ensure_ascii=False).encode("utf-8")
    return compile_schema(schema, lim).matches(text)


def test_exact_integer_ranges():
    for lo, hi in [(0, 9), (-3000, 70000), (123456, 123999), (-5, 5), (99, 100000), (-250, -17)]:
        dfa = compile_schema({"type": "integer", "minimum": lo, "maximum": hi}, LIM)
        for v in {lo - 1, lo, lo + 1, hi - 1, hi, hi + 1, 0, -1, (lo + hi) // 2, 10 * hi + 1}:
            assert dfa.matches(str(v).encode()) == (lo <= v <= hi), (lo, hi, v)
        assert not dfa.matches(b"007") and not dfa.matches(b"-0") and not dfa.matches(b"+1")
    one_sided = compile_schema({"type": "integer", "minimum": 5}, LIM)
    assert one_sided.matches(b"5") and one_sided.matches(b"999999999")
    assert not one_sided.matches(b"4") and not one_sided.matches(b"-7")
    below = compile_schema({"type": "integer", "exclusiveMaximum": 0}, LIM)
    assert below.matches(b"-1") and not below.matches(b"0") and not below.matches(b"3")


def test_exact_decimal_ranges():
    dfa = compile_schema({"type": "number", "minimum": 0.5, "maximum": 12.25}, LIM)
    inside = ["0.5", "0.50", "1", "12", "12.25", "12.2500", "3.1415", "0.9999", "7.0"]
    outside = ["0.4999", "12.2501", "13", "0", "-1", "12.3", "00.5", ".5", "1.", "1.23456"]
    assert all(dfa.matches(t.encode()) for t in inside)
    assert not any(dfa.matches(t.encode()) for t in outside)
    for t in inside:
        assert Decimal("0.5") <= Decimal(t) <= Decimal("12.25")
    crossing = compile_schema({"type": "number", "exclusiveMinimum": -1, "maximum": 1}, LIM)
    assert all(crossing.matches(t.encode()) for t in ["-0.9999", "0", "0.0", "1", "1.0000", "-0.5"])
    assert not any(crossing.matches(t.encode()) for t in ["-1", "-1.0", "1.0001", "-0", "-0.0", "2"])


def test_formats_and_patterns_accept_typical_instances():
    assert accepts({"type": "string", "format": "date"}, "2024-02-28")
    assert not accepts({"type": "string", "format": "date"}, "2024-13-01")
    assert accepts({"type": "string", "format": "date-time"}, "1999-12-01T23:59:59Z")
    assert accepts({"type": "string", "format": "uuid"}, str(uuid.UUID(int=0x1234, version=4)))
    assert accepts({"type": "string", "format": "ipv4"}, "192.168.0.255")
    assert not accepts({"type": "string", "format": "ipv4"}, "256.1.1.1")
    assert accepts({"type": "string", "format": "password"}, "hunter2")        # annotation only
    pat = {"type": "string", "pattern": r"^[a-z]{2,4}-\d+$"}
    assert accepts(pat, "ab-1") and accepts(pat, "wxyz-202") and not accepts(pat, "a-1")
    assert not accepts(pat, "abcde-1") and not accepts(pat, "ab-")
    unanchored = {"type": "string", "pattern": "ab+c"}
    assert accepts(unanchored, "xabbbcyy") and accepts(unanchored, "abc")
    assert not accepts(unanchored, "ac")
    quoted = {"type": "string", "pattern": r'^"[^"\\]{1,3}"$'}       # quotes need JSON escaping
    assert accepts(quoted, '"hi"') and not accepts(quoted, 'hi')
    newline = {"type": "string", "pattern": r"^a\sb$"}
    assert accepts(newline, "a\nb") and accepts(newline, "a b") and not accepts(newline, "ab")
    uni = {"type": "string", "pattern": r"^[^a-z]{2}$"}
    assert accepts(uni, "É€") and accepts(uni, "𝄞1") and not accepts(uni, "ab")
    both = {"type": "string", "pattern": r"^\w+$", "minLength": 3, "maxLength": 5}
    assert accepts(both, "abc") and accepts(both, "ab_12")
    assert not accepts(both, "ab") and not accepts(both, "abcdef") and not accepts(both, "a-c")


def test_pattern_outputs_match_the_python_regex():
    """Random members of the pattern automata re-checked with `re` on the decoded string."""
    rng = np.random.RandomState(5)
    for pattern in [r"^[a-z]{2,4}-\d+$", r"^(foo|ba[rz])+\.[A-Z]?$", r"x[^xy]{0,2}y", r"^\d{3}-\d{4}$",
                    r"^[\w.]+@[a-z]+\.(com|org)$", r"^\S+ \S+$"]:
        dfa = compile_schema({"type": "string", "pattern": pattern}, LIM)
        for _ in range(120):
            value = json.loads(random_accepted(dfa, rng).decode("utf-8"))
            assert re.search(pattern, value), (pattern, value)


def test_tuples_mappings_and_unique_items():
    tup = {"type": "array", "prefixItems": [{"type": "integer"}, {"type": "string"}],
           "items": False, "minItems": 2, "maxItems": 2}
    assert accepts(tup, [7, "x"]) and not accepts(tup, [7]) and not accepts(tup, [7, "x", 1])
    assert not accepts(tup, ["x", 7])
    open_tail = {"type": "array", "prefixItems": [{"type": "boolean"}], "items": {"type": "integer"},
                 "maxItems": 3}
    assert accepts(open_tail, [True]) and accepts(open_tail, [False, 1, 2])
    assert not accepts(open_tail, [True, 1, 2, 3]) and not accepts(open_tail, [1])
    assert accepts(open_tail, [])                                        # minItems defaults to 0
    mapping = {"type": "object", "additionalProperties": {"type": "integer"}, "minProperties": 1}
    assert accepts(mapping, {"a": 1}) and accepts(mapping, {"a": 1, "bc": -2})
    assert not accepts(mapping, {}) and not accepts(mapping, {"a": "x"})
    keyed = {"type": "object", "additionalProperties": {"type": "boolean"},
             "propertyNames": {"pattern": "^k[0-9]$"}}
    assert accepts(keyed, {"k1": True, "k2": False}) and not accepts(keyed, {"x": True})
    uniq = {"type": "array", "items": {"enum": ["a", "b", "c"]}, "uniqueItems": True, "minItems": 1}
    assert accepts(uniq, ["b", "a"]) and accepts(uniq, ["c"]) and not accepts(uniq, ["a", "a"])
    assert not accepts(uniq, [])


@pytest.mark.parametrize("schema", [
    {"type": "string", "format": "hostname"},
    {"type": "string", "pattern": r"\bfoo"},
    {"type": "string", "pattern": r"(a)\1"},
    {"type": "string", "pattern": r"(?i)abc"},
    {"type": "string", "pattern": "("},
    {"type": "integer", "multipleOf": 3},
    {"type": "number", "multipleOf": 0.001, "minimum": 0, "maximum": 1000},
    {"type": "array", "items": {"type": "integer"}, "uniqueItems": True},
    {"type": "array", "items": {"type": "integer"}, "contains": {"const": 1}},
    {"type": "object", "properties": {"a": {"type": "integer"}}, "patternProperties": {"^x": {}}},
    {"type": "object", "properties": {"a": {"type": "integer"}}, "dependentRequired": {"a": ["b"]}},
    {"not": {"type": "string"}},
    {"type": "string", "minLength": 3, "maxLength": 2},
    {"type": "integer", "minimum": 3, "maximum": 2},
    {"type": "string", "enum": []},
    {"allOf": [{"type": "integer", "minimum": 0}, {"type": "integer", "minimum": 5}]},
    {"type": "string", "frobnicate": 1},
    {"type": "string", "pattern": "^a{3}$", "maxLength": 2},
])
def test_unsupported_or_contradictory_schemas_raise(schema):
    with pytest.raises(SchemaError):
        compile_schema(schema, LIM)
    assert issubclass(SchemaError, ValueError)         # the SDK's argument-error convention


def test_allof_merges_compatible_parts_and_annotations_are_ignored():
    merged = {"allOf": [{"type": "integer", "minimum": 2}, {"maximum": 4}], "title": "T",
              "description": "d", "default": 3}
    dfa = compile_schema(merged, LIM)
    assert [dfa.matches(str(v).encode()) for v in (1, 2, 4, 5)] == [False, True, True, False]
    assert compile_schema({"title": "Anything"}, LIM).matches(b'"x"')
    assert compile_schema({"title": "Anything"}, LIM).matches(b"null")


def test_pathological_patterns_are_refused_not_hung():
    import time
    t0 = time.perf_counter()
    with pytest.raises(SchemaError):
        compile_schema({"type": "string", "pattern": r"^(a|b)*a(a|b){24}$"}, FsmLimits())
    assert time.perf_counter() - t0 < 120
    # the common unanchored shape stays small with the default 64-character cap
    dfa = compile_schema({"type": "string", "pattern": r"ab+c"}, FsmLimits())
    assert dfa.n_states < 10000 and dfa.matches(b'"' + b"x" * 30 + b"abbc" + b"y" * 30 + b'"')
    assert not dfa.matches(b'"' + b"x" * 40 + b"abbc" + b"y" * 30 + b'"')


def test_unions_keep_the_alternatives_that_can_be_expressed():
    """pydantic's Decimal is number | string-with-a-lookahead-pattern: the number form is kept,
    so the field still works; a union with no expressible alternative raises."""
    import ipaddress
    import pathlib
    from decimal import Decimal as D

    class Money(BaseModel):
        amount: D
        wait: datetime.timedelta
        host: ipaddress.IPv4Address
        where: pathlib.Path

    dfa = compile_schema(Money.model_json_schema(), LIM)
    rng = np.random.RandomState(3)
    for _ in range(150):
        Money.model_validate_json(random_accepted(dfa, rng))
    with pytest.raises(SchemaError):
        compile_schema({"anyOf": [{"type": "string", "pattern": r"(?=a)b"}, {"not": {}}]}, LIM)


def test_recursive_models_are_unrolled_to_a_fixed_depth():
    class Node(BaseModel):
        value: int = Field(ge=0, le=9)
        children: List["Node"] = Field(default_factory=list, max_length=2)

    class Linked(BaseModel):
        tag: Literal["x", "y"]
        next: Optional["Linked"] = None

    for model in (Node, Linked):
        schema = model.model_json_schema()
        dfa = compile_schema(schema, LIM)
        rng = np.random.RandomState(2)
        depths = set()
        for _ in range(200):
            text = random_accepted(dfa, rng)
            model.model_validate_json(text)
            depths.add(text.count(b"[") + text.count(b"{") - 1)
        assert max(depths) >= 2                      # nesting really is produced
    leaf = {"value": 1, "children": []}
    mid = {"value": 2, "children": [leaf, leaf]}
    assert accepts(Node.model_json_schema(), {"value": 3, "children": [mid]})
    too_deep = {"value": 0, "children": [{"value": 0, "children": [{"value": 0, "children": [leaf]}]}]}
    assert not accepts(Node.model_json_schema(), too_deep)
    with pytest.raises(SchemaError):                 # a cycle with no way to stop
        compile_schema({"$defs": {"A": {"type": "object", "properties": {"a": {"$ref": "#/$defs/A"}},
                                        "required": ["a"]}}, "$ref": "#/$defs/A"}, LIM)


def test_type_list_with_keywords_of_one_member():
    """{"type": ["string", "null"], "maxLength": 5}: maxLength applies to the string branch
    only (it is vacuous for null), the schema compiles, both branches are accepted."""
    from sutro_b200.schema_fsm import compile_schema
    d = compile_schema({"type": ["string", "null"], "maxLength": 5})
    assert d.matches(b"null") and d.matches(b'"abcde"') and not d.matches(b'"abcdef"')
    d = compile_schema({"type": ["integer", "string"], "minimum": 3, "maximum": 9, "maxLength": 1})
    assert d.matches(b"7") and not d.matches(b"2") and d.matches(b'"x"') and not d.matches(b'"xy"')


def test_longest_path_bounds_every_accepted_string():
    from sutro_b200.schema_fsm import FsmLimits, compile_schema
    d = compile_schema({"type": "object", "properties": {"sentiment": {"type": "string", "enum": [
        "positive", "neutral", "negative"]}}, "required": ["sentiment"]})
    assert d.longest_path() == len('{"sentiment":"positive"}')
    d = compile_schema({"type": "array", "items": {"type": "integer", "minimum": 0, "maximum": 99},
                        "maxItems": 3})
    assert d.longest_path() == len("[99,99,99]")
    d = compile_schema({"type": "string"}, FsmLimits(max_string_chars=4))
    assert d.longest_path() == 2 + 4 * 6        # four \\uXXXX escapes between the quotes


def test_thinking_turn_automaton():
    """`<model>-thinking`: free reasoning (capped, no '<'), the closing </think> line, then the
    schema instance or free text (reference output shape: sutro/sdk.py:1155-1164)."""
    from sutro_b200.schema_fsm import compile_thinking, split_thinking
    sch = {"type": "object", "properties": {"sentiment": {"type": "string", "enum": ["a", "b"]}}}
    d = compile_thinking(sch, None, 16)
    assert d.matches(b'hm, ok\n</think>\n\n{"sentiment":"a"}')
    assert d.matches(b'</think>\n\n{"sentiment":"b"}')                     # empty reasoning
    assert not d.matches(b'{"sentiment":"a"}')                             # the block is not optional
    assert not d.matches(b'x<y</think>\n\n{"sentiment":"a"}')              # '<' only opens the close tag
    assert not d.matches(b'12345678901234567</think>\n\n{"sentiment":"a"}')   # 17 > 16 characters
    assert d.longest_path() == 16 * 4 + len("</think>\n\n") + len('{"sentiment":"a"}')
    free = compile_thinking(None, None, 8)
    assert free.matches("é ok</think>\n\nfree <text> here".encode()) and free.longest_path() is None
    assert split_thinking('hm, ok\n</think>\n\n{"sentiment":"a"}') == ("hm, ok", '{"sentiment":"a"}')
    assert split_thinking("never closed") == ("never closed", "")
