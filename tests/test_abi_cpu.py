"""The C-ABI shared library loads without a GPU and exports every symbol that
include/sutro_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "sutro_b200.h")
LIB = os.path.join(ROOT, "sutro_b200", "libsutro_b200.so")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    names = set(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", text))
    names -= {"sb200_progress_fn"}
    return sorted(names)


def test_header_declares_the_expected_surface():
    names = declared_symbols()
    for must in ["sb200_engine_create", "sb200_engine_run", "sb200_tokenizer_encode",
                 "sb200_tokenizer_decode", "sb200_gemm_bf16_tn", "sb200_attn_decode",
                 "sb200_attn_prefill", "sb200_rope_kv_write", "sb200_rmsnorm",
                 "sb200_fsm_build_mask", "sb200_last_error"]:
        assert must in names


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built (run __graft_entry__.build())")
def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(LIB)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, missing
    lib.sb200_abi_version.restype = ctypes.c_int
    assert lib.sb200_abi_version() == 1
    lib.sb200_last_error.restype = ctypes.c_char_p
    assert lib.sb200_last_error() is not None


@pytest.mark.skipif(not os.path.exists(LIB), reason="library not built")
def test_python_binding_table_matches_header():
    from sutro_b200 import _lib, engine  # noqa: F401  (engine registers its entry points)
    assert set(_lib.exported_symbols()) == set(declared_symbols())
    _lib.lib()


def test_no_product_module_imports_the_oracle():
    pkg = os.path.join(ROOT, "sutro_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f


def test_ctypes_struct_layouts_match_the_header(tmp_path):
    """The Python mirrors of the C structs (engine.py) against the header itself: a probe
    compiled with gcc prints sizeof/offsetof for every field."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    from sutro_b200 import engine as E
    mirrors = {"sb200_engine_config": E.EngineConfigC, "sb200_engine_weights": E.EngineWeightsC,
               "sb200_job": E.JobC, "sb200_job_stats": E.JobStatsC, "sb200_result": E.ResultC}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{HEADER}"', "int main(void){"]
    for cname, mirror in mirrors.items():
        lines.append(f'printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, *_ in mirror._fields_:
            lines.append(f'printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0;}")
    src = tmp_path / "probe.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "probe"
    subprocess.run(["gcc", "-std=c11", "-o", str(exe), str(src)], check=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    seen = 0
    for line in out.splitlines():
        cname, what, value = line.split()
        mirror = mirrors[cname]
        want = ctypes.sizeof(mirror) if what == "size" else getattr(mirror, what).offset
        assert int(value) == want, (cname, what, value, want)
        seen += 1
    assert seen == sum(len(m._fields_) + 1 for m in mirrors.values())
