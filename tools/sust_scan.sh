#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "gemm or qkv_rope" -p no:cacheprovider 2>&1 | tail -3
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
timeout 200 python tools/gemm_sustained.py 32768 --secs 2
timeout 500 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/c18_bench.json 2> gpurun_out/c18_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c18_bench.json"))
print(round(d["value"], 1), "rows/s e2e", round(d["e2e"]["value"], 1),
      {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items() if v > 1}, "gemm TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"])
PY
