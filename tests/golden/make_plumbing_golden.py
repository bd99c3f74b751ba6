"""Golden vectors for the host-side plumbing, produced by the REFERENCE's own functions.

Run in the build container (needs /root/reference; the GPU box never runs this):

    python tests/golden/make_plumbing_golden.py

The reference package cannot be imported whole (polars / colorama / yaspin are not installed
and there is no network), so this script loads two of its source files by path with empty
stand-in modules for the missing imports.  Only list / pandas code paths are exercised — the
stand-ins are never called.  Output: tests/golden/plumbing.json, consumed by
tests/test_golden.py against sutro_b200.common / sutro_b200.templates.
"""
import importlib.util
import json
import os
import sys
import tempfile
import types

import pandas as pd
from pydantic import BaseModel

REF = "/root/reference/sutro"
HERE = os.path.dirname(os.path.abspath(__file__))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def load_reference():
    class _NoFrame:      # isinstance(x, pl.DataFrame) must be False for pandas/list input
        pass
    if "polars" not in sys.modules:
        _stub("polars", DataFrame=_NoFrame)
    if "colorama" not in sys.modules:
        _stub("colorama", Fore=types.SimpleNamespace(), Style=types.SimpleNamespace())
    pkg = _stub("sutro")
    pkg.__path__ = [REF]
    tp = _stub("sutro.templates")
    tp.__path__ = [os.path.join(REF, "templates")]

    def load(mod, rel):
        spec = importlib.util.spec_from_file_location(mod, os.path.join(REF, rel))
        m = importlib.util.module_from_spec(spec)
        sys.modules[mod] = m
        spec.loader.exec_module(m)
        return m
    common = load("sutro.common", "common.py")
    load("sutro.interfaces", "interfaces.py")
    evals = load("sutro.templates.evals", "templates/evals.py")
    return common, evals


class Sentiment(BaseModel):
    sentiment: str


class Item(BaseModel):
    name: str
    qty: int


class Order(BaseModel):
    items: list[Item]
    note: str | None = None


def frames():
    f1 = pd.DataFrame({"title": ["a", None, "c"], "body": ["x", "y", None], "n": [1, 2, 3]})
    f2 = pd.DataFrame({"q": ["why?", "how?"], "ctx": ["because", ""], "score": [0.5, None]})
    return {"f1": f1, "f2": f2}


CONCAT_CASES = [("f1", ["title", ": ", "body"]), ("f1", ["n", "-", "title", "-", "n"]),
                ("f1", ["body"]), ("f2", ["Q: ", "q", "\nC: ", "ctx", " (", "score", ")"]),
                ("f2", ["ctx", "ctx"])]

ELO_CASES = [
    dict(ballots=[["B", "A", "C"], ["A", "B", "C"], ["B", "C", "A"], ["B", "A", "C"]]),
    dict(ballots=[["x", "y"], ["x", "y"], ["y", "x"]], laplace=0.0),
    dict(ballots=[["B", ["A", "C"], "D"], ["D", "B", "A", "C"], [["A", "B"], "C", "D"]]),
    dict(ballots=[["m1", None, "m2", "m3"], ["m3", "m1", "m2"], ["m2", "m3", "m1"]] * 3,
         laplace=1.0, elo_mean=1000.0),
    dict(ballots=[["a", "b", "c", "d", "e"]] * 5 + [["e", "d", "c", "b", "a"]] * 2, max_iter=50),
    dict(ballots=[[1, 2, 3], [2, 1, 3], [3, 1, 2]], tol=1e-12),
]


def main():
    common, evals = load_reference()
    out = {"source": "sutro/common.py:72-163 and sutro/templates/evals.py:182-334 executed from "
                     "/root/reference", "concat": [], "handle": [], "schema": [], "elo": []}
    fr = frames()
    for fname, cols in CONCAT_CASES:
        out["concat"].append({"frame": fname, "column": cols,
                              "expect": common.do_dataframe_column_concatenation(fr[fname], cols)})
    out["handle"].append({"kind": "list", "expect": common.handle_data_helper(["p", "q"])})
    out["handle"].append({"kind": "frame_str", "frame": "f2", "column": "q",
                          "expect": common.handle_data_helper(fr["f2"], "q")})
    out["handle"].append({"kind": "frame_list", "frame": "f1", "column": ["title", "/", "n"],
                          "expect": common.handle_data_helper(fr["f1"], ["title", "/", "n"])})
    out["handle"].append({"kind": "dataset", "data": "dataset-abc", "column": "text",
                          "expect": common.handle_data_helper("dataset-abc", "text")})
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False) as f:
        f.write("  first line \nsecond\n\n third\t\n")
    out["handle"].append({"kind": "txt", "content": "  first line \nsecond\n\n third\t\n",
                          "expect": common.handle_data_helper(f.name)})
    os.unlink(f.name)
    for kind, bad in (("frame_no_column", (fr["f1"], None)), ("bad_type", (42, None)),
                      ("bad_ext", ("/tmp/x.json", "c"))):
        try:
            common.handle_data_helper(*bad)
            err = None
        except ValueError as e:
            err = str(e)
        out["handle"].append({"kind": kind, "error": err})
    for nm, sch in (("Sentiment", Sentiment), ("Order", Order)):
        out["schema"].append({"model": nm, "expect": common.normalize_output_schema(sch)})
    out["schema"].append({"model": "dict", "expect": common.normalize_output_schema({"type": "object"})})
    try:
        common.normalize_output_schema("nope")
    except ValueError as e:
        out["schema"].append({"model": "str", "error": str(e)})
    for case in ELO_CASES:
        kw = {k: v for k, v in case.items() if k != "ballots"}
        tbl = evals.Rank.elo(case["ballots"], **kw)
        out["elo"].append({"ballots": case["ballots"], "kwargs": kw,
                           "index": [str(i) for i in tbl.index],
                           "columns": {c: [float(x) for x in tbl[c]] for c in tbl.columns}})
    with open(os.path.join(HERE, "plumbing.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote plumbing.json:", {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    main()
