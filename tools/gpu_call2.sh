#!/usr/bin/env bash
# round 2, GPU call 2: whole GPU suite (new parity / public-API / real-size tests), then bench N=1
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_flips.jsonl gpurun_out/parity_real_size.jsonl
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider 2>&1 | tail -40 > gpurun_out/c2_pytest.txt
tail -25 gpurun_out/c2_pytest.txt
timeout 900 python bench.py --steps 2 --warmup 2 > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/c2_bench.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c2_bench.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "gpu_launches")}, d["e2e"]["value"])
    print("kernel_ms", {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items()})
    print("roofline", round(d["roofline"]["frac"], 3), "cpu", d["cpu_baseline"] and d["cpu_baseline"].get("outputs_equal"))
    print("shape", d["config"]["shape"])
    for k, v in (d.get("secondary") or {}).items():
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if not isinstance(b, dict)})
except Exception as e:
    print("bench parse failed", e)
PY
