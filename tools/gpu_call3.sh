#!/usr/bin/env bash
set -u
mkdir -p gpurun_out
rm -f gpurun_out/parity_real_size.jsonl
# the rewritten softmax branch of the tcgen05 attention first: stop here if it is wrong
timeout -k 10 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dense" -p no:cacheprovider 2>&1 | tail -8
if [ ${PIPESTATUS[0]} -ne 0 ]; then echo "DENSE ATTENTION TESTS FAILED - stopping"; exit 1; fi
timeout 120 python tools/attn_bench.py 2>&1 | tail -4
timeout 120 python tools/attn_bench.py --new 512 --past 0 2>&1 | tail -3
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_zy_real_size_parity_gpu.py tests/test_c_host_gpu.py tests/test_public_api_gpu.py -q --timeout 900 -p no:cacheprovider \
   -k "tile_configurations or row_order or real_size or c_program or public or thinking or templates" 2>&1 | tail -80 > gpurun_out/c3_pytest.txt
tail -70 gpurun_out/c3_pytest.txt
SB200_QKV_DENSE=1 timeout 120 python tools/qkv_epi_bench.py 2>&1 | tail -3
SB200_QKV_DENSE=0 timeout 120 python tools/qkv_epi_bench.py 2>&1 | head -1
timeout 200 python tools/gemm_bench.py 32768 2>&1 | grep -v "variant=51[46]" | tail -12
# representative ncu capture of the tcgen05 prefill attention (one launch, the bench shape)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_prefill_tc -s 3 -c 1 \
   -o gpurun_out/r02_attn_prefill_tc python tools/attn_bench.py --iters 1 --new 110 > gpurun_out/c3_ncu.log 2>&1
tail -2 gpurun_out/c3_ncu.log
for tc in 1 0; do
  SB200_PREFILL_TC=$tc timeout 400 python bench.py --steps 1 --warmup 1 --no-secondary --no-cpu-baseline \
     > gpurun_out/c3_bench_tc$tc.json 2> gpurun_out/c3_bench_tc$tc.err
  python - $tc <<'PY'
import json, sys
d = json.load(open(f"gpurun_out/c3_bench_tc{sys.argv[1]}.json"))
print("PREFILL_TC=" + sys.argv[1], round(d["value"], 1), "rows/s", {k: round(v, 1) for k, v in d["kernel_ms_profiled_job"].items()},
      "gemm TF/s", round(d["roofline"]["achieved"], 1), d["clocks"]["sm_mhz"])
PY
done
