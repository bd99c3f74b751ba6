#!/usr/bin/env bash
# Round-2 first step for prefill attention variant 2 (csrc/attn_prefill_v2.cu, opt-in):
# parity under the env switch, then the same short benchmark with and without it.
#   gpurun --timeout 600 -- 'bash tools/try_prefill_v2.sh'
set -u
mkdir -p gpurun_out
# pretrained-loading path (tests/test_engine_gpu.py::test_model_directory_on_disk_matches_oracle)
SB200_TEST_PRETRAINED=1 timeout 200 python -m pytest tests/test_engine_gpu.py -x -q -k model_directory \
    2>&1 | tail -5
SB200_PREFILL_V2=1 timeout 300 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q \
    2>&1 | tail -5
for v in 0 1; do
  SB200_PREFILL_V2=$v timeout 200 python bench.py --rows 8000 --steps 2 --warmup 3 --no-cpu-baseline \
      2> gpurun_out/pfv2_$v.err > gpurun_out/pfv2_$v.json
  python - "$v" <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/pfv2_{sys.argv[1]}.json"))
print("SB200_PREFILL_V2=" + sys.argv[1], round(d["value"], 1), "rows/s",
      {k: round(v, 1) for k, v in d["kernel_ms_profiled_step"].items()}, d["clocks"]["sm_mhz"])
PY
done
