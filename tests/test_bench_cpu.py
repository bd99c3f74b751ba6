"""bench.py is the file the driver runs unattended: check on CPU that it imports, parses its
command line, and that its host-only helpers work (the GPU arms need a B200)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_bench_imports_and_prints_usage():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--help"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    for flag in ("--gpus", "--steps", "--warmup", "--impl", "--rows", "--workload"):
        assert flag in out.stdout


def test_host_helpers():
    sys.path.insert(0, ROOT)
    import bench
    from sutro_b200 import synth
    rows = synth.product_reviews(500, seed=0)
    p = bench.plumbing_cost("qwen-3-4b", rows, ['{"sentiment":"positive"}'] * len(rows))
    assert p["rows"] == 500 and p["rows_per_s"] > 0 and p["payload_bytes"] > 500
    t = bench.ncu_traffic("r01_ncu_gemm_gateup_raw.csv")
    assert t is None or t > 1e6                      # bytes of one launch, when the capture exists
    assert bench.ncu_traffic("does_not_exist.csv") is None
    pk = bench.peaks()
    assert pk["hbm_gbs"] > 1000 and pk["bf16_tflops_sustained"] > 100
    json.dumps(bench.SCHEMA)
    assert len(bench.make_rows(7, seed=1)) == 7


def test_smoke_and_build_entry_points_exist():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    assert callable(g.build) and callable(g.smoke)
