#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dense" -p no:cacheprovider 2>&1 | tail -6
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "DENSE ATTENTION TESTS FAILED - stopping"; exit 1; fi
timeout 120 python tools/attn_bench.py 2>&1 | tail -4
timeout 120 python tools/attn_bench.py --new 512 --past 0 2>&1 | tail -3
timeout 600 python -m pytest tests/test_engine_gpu.py -q --timeout 600 -p no:cacheprovider -k "row_order or greedy or geometry" 2>&1 | tail -8
for cfg in "1 0" "1 256" "0 0"; do
  set -- $cfg
  SB200_PREFILL_TC=$1 SB200_GEMM_SWIGLU_BN=$2 timeout 500 python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline \
     > gpurun_out/c4_bench_$1_$2.json 2> gpurun_out/c4_bench_$1_$2.err
  python - $1 $2 <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/c4_bench_{sys.argv[1]}_{sys.argv[2]}.json"))
print(f"TC={sys.argv[1]} SWIGLU_BN={sys.argv[2]}", round(d["value"], 1), "rows/s e2e", round(d["e2e"]["value"], 1),
      {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items() if v > 1}, "gemm TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"])
PY
done
