"""Host-side schema -> DFA compiler: every accepted string is valid JSON that validates
against the schema; compact serialisations of valid instances are accepted."""
import json
from typing import List, Literal, Optional

import numpy as np
import pytest
from pydantic import BaseModel, Field

from oracle.fsm_ref import TokenFSM, random_accepted, validates
from sutro_b200 import vocab as V
from sutro_b200.schema_fsm import FsmLimits, SchemaError, compile_schema


class Sentiment(BaseModel):            # README.md:45-46 of the reference
    sentiment: str


class SentimentEnum(BaseModel):
    sentiment: Literal["positive", "neutral", "negative"]


class Classification(BaseModel):       # templates/classification.py:58-89 shape
    scratchpad: str = Field(max_length=40)
    classification: str = Field(max_length=16)


class Score(BaseModel):                # templates/evals.py:42-52 shape
    score: int = Field(ge=1, le=10)


class Item(BaseModel):
    name: str = Field(max_length=12)
    quantity: int = Field(ge=0, le=1000)
    kind: Literal["a", "b", "c"]
    price: Optional[float] = None


class Order(BaseModel):                # nested, $defs/$ref, arrays (config 5)
    customer: str = Field(max_length=10)
    items: List[Item] = Field(max_length=3)
    paid: bool


SCHEMAS = [Sentiment, SentimentEnum, Classification, Score, Order]
LIM = FsmLimits(max_string_chars=12, max_array_items=3)


@pytest.mark.parametrize("model", SCHEMAS)
def test_accepted_strings_parse_and_validate(model):
    schema = model.model_json_schema()
    dfa = compile_schema(schema, LIM)
    rng = np.random.RandomState(0)
    for _ in range(150):
        s = random_accepted(dfa, rng)
        obj = json.loads(s.decode("utf-8"))          # valid UTF-8 + valid JSON
        assert validates(obj, schema), s
        model.model_validate(obj)                    # pydantic agrees


def test_valid_instances_are_accepted_and_invalid_rejected():
    dfa = compile_schema(Order.model_json_schema(), LIM)
    good = Order(customer="Ann", items=[Item(name="x", quantity=3, kind="b", price=1.5),
                                        Item(name="yé", quantity=0, kind="a")], paid=True)
    text = good.model_dump_json()
    assert dfa.matches(text.encode())
    assert not dfa.matches(text.replace('"b"', '"z"').encode())       # enum violation
    assert not dfa.matches(text.replace("true", "1").encode())        # type violation
    assert not dfa.matches(text[:-1].encode())                        # truncated
    assert not dfa.matches(text.replace('"Ann"', '"' + "A" * 11 + '"').encode())  # maxLength
    assert dfa.matches(Order(customer="", items=[], paid=False).model_dump_json().encode())

    d2 = compile_schema(Score.model_json_schema(), LIM)
    for v in range(-2, 14):
        assert d2.matches(json.dumps({"score": v}, separators=(",", ":")).encode()) == (1 <= v <= 10)

    d3 = compile_schema({"type": "object", "properties": {"s": {"type": "string", "enum":
                         ["positive", "negative"]}}}, LIM)   # reference tests/test_sdk.py:427-435
    assert d3.matches(b'{"s":"positive"}') and not d3.matches(b'{"s":"neutral"}')


def test_string_escapes_and_utf8():
    dfa = compile_schema({"type": "string", "maxLength": 6}, LIM)
    assert dfa.matches('"a\\"b"'.encode()) and dfa.matches('"\\u00e9x"'.encode())
    assert dfa.matches('"日本語"'.encode()) and dfa.matches('"🙂"'.encode())
    assert not dfa.matches(b'"\xff"') and not dfa.matches(b'"\xe3\x81"')   # malformed UTF-8
    assert not dfa.matches(b'"a\nb"')                                      # raw control char
    assert not dfa.matches('"1234567"'.encode())                           # 7 chars > 6


def test_unsupported_schema_raises_value_error():
    with pytest.raises(ValueError):      # a cycle nothing can stop (a required self-reference)
        compile_schema({"type": "object", "properties": {"x": {"$ref": "#/$defs/T"}},
                        "$defs": {"T": {"type": "object", "properties": {"t": {"$ref": "#/$defs/T"}},
                                        "required": ["t"]}}})
    # a cycle through an array is unrolled to a fixed depth and bottoms out in []
    nested = compile_schema({"type": "object", "properties": {"x": {"$ref": "#/$defs/T"}},
                             "$defs": {"T": {"type": "array", "items": {"$ref": "#/$defs/T"},
                                             "maxItems": 2}}})
    assert nested.matches(b'{"x":[[],[[]]]}') and not nested.matches(b'{"x":[[[[[]]]]]}')
    with pytest.raises(SchemaError):
        compile_schema({"type": "frobnicate"})


def test_token_fsm_only_allows_live_tokens_and_terminates():
    v = V.build_vocab("qwen3", 2048, seed=0, n_trained=600)
    dfa = compile_schema(SentimentEnum.model_json_schema(), LIM)
    fsm = TokenFSM(dfa, v)
    rng = np.random.RandomState(1)
    for _ in range(20):
        s, out = fsm.start, []
        for _ in range(200):
            ok = fsm.allowed(s).numpy()
            assert ok.any()
            tok = int(rng.choice(np.nonzero(ok)[0]))
            assert tok != v.eos_id or dfa.accept[s]
            if tok == v.eos_id:
                break
            out.append(tok)
            s = fsm.step(s, tok)
            assert s >= 0
            if fsm.is_final(s):
                break
        obj = json.loads(v.decode(out).decode())
        assert obj["sentiment"] in ("positive", "neutral", "negative")
