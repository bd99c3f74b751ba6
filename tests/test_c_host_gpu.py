"""The boundary without Python: a C program (tests/c_host/infer_rows.c, compiled here with gcc
and linked against libsutro_b200.so only) opens a model bundle, hands over the system prompt,
the JSON schema TEXT and three rows, and must print exactly what the Python host produces
(LocalEngine.generate) — template rendering, schema compilation, jump-forward plan and output
budget all happen behind the C-ABI (csrc/model_bundle.cu, csrc/schema_compile.cu)."""
import json
import os
import shutil
import subprocess
from typing import Literal

import pytest
from pydantic import BaseModel

from sutro_b200 import modelspec as MS
from sutro_b200 import synth, vocab as VB

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class Sentiment(BaseModel):
    sentiment: Literal["positive", "neutral", "negative"]


def _compile(tmp_path):
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = tmp_path / "infer_rows"
    subprocess.run(["gcc", "-std=c11", "-O1", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c_host", "infer_rows.c"),
                    "-L", os.path.join(ROOT, "sutro_b200"), "-lsutro_b200",
                    "-Wl,-rpath," + os.path.join(ROOT, "sutro_b200"), "-o", str(exe)], check=True)
    return str(exe)


@pytest.mark.parametrize("name", ["tiny-qwen3", "tiny-llama"])
def test_c_program_matches_the_python_host(name, tmp_path):
    from sutro_b200.bundle import export_bundle
    from sutro_b200.engine import LocalEngine
    exe = _compile(tmp_path)
    spec = MS.get_spec(name)
    w = MS.make_weights(spec, seed=0, std=0.05)
    v = VB.build_vocab(spec.family, spec.vocab_size, seed=0, n_trained=600)
    eng = LocalEngine(spec, MS.pack_for_engine(spec, w, "cuda"), v, device=0, kv_pages=512,
                      max_slots=8, max_prefill_tokens=512)
    bundle = export_bundle(eng, str(tmp_path / "bundle"))
    schema = Sentiment.model_json_schema()
    (tmp_path / "schema.json").write_text(json.dumps(schema))
    rows = synth.README_REVIEWS + ["", "café naïve 東京"]
    want = eng.generate(rows, system_prompt=synth.README_SYSTEM_PROMPT, json_schema=schema,
                        max_new_tokens=len('{"sentiment":"positive"}')).outputs
    free = eng.generate(rows, system_prompt=None, max_new_tokens=256).outputs
    eng.close()
    out = subprocess.run([exe, bundle, synth.README_SYSTEM_PROMPT, str(tmp_path / "schema.json")] + rows,
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = out.stdout.split("\n")[:len(rows)]
    assert got == want, (got, want, out.stderr[-500:])
    for text in got:
        Sentiment.model_validate(json.loads(text))
    # unconstrained, no system prompt: the default budget (512, capped at half the window)
    out = subprocess.run([exe, bundle, "", "-"] + rows, capture_output=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = out.stdout.decode("utf-8", errors="replace").split("\n")
    assert len(got) >= len(rows)
    assert got[0] == free[0].split("\n")[0]
    # an unsupported schema keyword is an argument error, reported, never ignored
    (tmp_path / "bad.json").write_text(json.dumps({"type": "string", "pattern": "^a+$"}))
    out = subprocess.run([exe, bundle, "x", str(tmp_path / "bad.json"), "row"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 1 and "pattern" in out.stderr
