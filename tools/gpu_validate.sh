#!/usr/bin/env bash
# Final validation, part 1: whole GPU suite, smoke, attention timeline, default bench (N=1).
set -u
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --timeout 900 -p no:cacheprovider > gpurun_out/f1_pytest.txt 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/f1_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 120 python tools/attn_bench.py > gpurun_out/f1_attn_bench.txt 2>&1; cat gpurun_out/f1_attn_bench.txt | tail -4
SB200_LIB=$PWD/tools/_trace/libsutro_b200_trace.so timeout 120 python tools/attn_bench.py --trace 0 --trace-from 40 --trace-n 12 > gpurun_out/f1_attn_trace.txt 2>&1; tail -3 gpurun_out/f1_attn_trace.txt
timeout 1500 python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/f1_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/f1_bench.json"))
print(round(d["value"], 1), "rows/s e2e", round(d["e2e"]["value"], 1),
      {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items() if v > 1}, "gemm TF/s", round(d["roofline"]["achieved"], 1),
      "frac", round(d["roofline"]["frac"], 3), d["clocks"]["sm_mhz"], d.get("cpu_baseline", {}))
for k, v in (d.get("secondary") or {}).items():
    print(k, {kk: (round(vv, 1) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("rows_per_sec", "value", "output_tokens_per_sec", "attn_decode_gbs_non_shared_kv", "attn_decode_share_of_step", "valid_json")})
PY
