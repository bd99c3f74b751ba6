#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_kernels_gpu.py -x -q -k "dense or qkv_rope or rope_kv" -p no:cacheprovider 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "KERNEL TESTS FAILED - stopping"; exit 1; fi
timeout 120 python tools/attn_bench.py 2>&1 | tail -4
timeout 120 python tools/attn_bench.py --new 512 --past 0 2>&1 | tail -3
SB200_QKV_DENSE=1 timeout 120 python tools/qkv_epi_bench.py 2>&1 | tail -3
for g in auto 1 2 3 4 6 9 16 32; do
  if [ $g = auto ]; then timeout 120 python tools/gemm_raster_scan.py; else SB200_GEMM_GM=$g timeout 120 python tools/gemm_raster_scan.py; fi
done
timeout 900 python -m pytest tests/test_engine_gpu.py -q --timeout 600 -p no:cacheprovider 2>&1 | tail -6
timeout 500 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline > gpurun_out/c9_bench.json 2> gpurun_out/c9_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/c9_bench.json"))
print(round(d["value"], 1), "rows/s e2e", round(d["e2e"]["value"], 1),
      {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items() if v > 1}, "gemm TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"])
PY
