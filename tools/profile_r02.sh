#!/bin/bash
# Round-2 ncu evidence (runs on the GPU box via gpurun; outputs in gpurun_out/, summaries are
# extracted and committed under profiles/ by tools/profile_extract.py).
#   1. launch list of a small headline job (per-launch gpu__time_duration)
#   2. --set full captures, ONE app run: after the (unrepresentative, 46-token) prefix prefill,
#      the first launches of a full 32k-token prefill step: RMSNorm, fused-QKV GEMM, tcgen05
#      prefill attention, O-proj GEMM, RMSNorm, gate/up GEMM (SwiGLU), down GEMM
#   3. --set full captures of the decode side on the configs[2] shape (512-token documents,
#      B = 512): tokenizer kernels, sampler, decode attention
set -x
mkdir -p gpurun_out
R=${ROUND:-r02}
HEAD="python bench.py --rows 2048 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --kv-pages 16384"
DOCS="python bench.py --workload docs --rows 512 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --kv-pages 32768 --max-slots 512"
NCU="ncu --clock-control none --kernel-name-base demangled"
timeout 600 $NCU --metrics gpu__time_duration.sum -c 30000 --csv --log-file gpurun_out/launches_${R}.csv \
    $HEAD > gpurun_out/ncu_launch_${R}.log 2>&1
timeout 900 $NCU --set full --import-source on \
    -k 'regex:gemm2_bf16_tn_kernel|gemm_bf16_tn_kernel|attn_prefill_tc_kernel|rmsnorm_kernel' -s 256 -c 14 \
    -f -o gpurun_out/prof_prefill_step_${R} $HEAD > gpurun_out/ncu_prefill_step_${R}.log 2>&1
timeout 900 $NCU --set full --import-source on \
    -k 'regex:attn_decode|sample_greedy|bpe_kernel|pretok_kernel|detok|compact_rows|fsm_build' -c 44 \
    -f -o gpurun_out/prof_decode_side_${R} $DOCS > gpurun_out/ncu_decode_side_${R}.log 2>&1
tail -3 gpurun_out/ncu_prefill_step_${R}.log gpurun_out/ncu_decode_side_${R}.log
ls -la gpurun_out | tail -8
