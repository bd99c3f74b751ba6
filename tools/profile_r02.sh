#!/bin/bash
# Round-2 ncu evidence (runs on the GPU box via gpurun).  gpurun copies back at most 64 MiB, so
# the raw pages are exported to CSV on the box and the big .ncu-rep files are deleted; one
# small report with source (the down-projection GEMM + the tcgen05 attention) is kept.
#   1. launch list of a small headline job (per-launch gpu__time_duration)
#   2. --set full, ONE app run: after the (46-token) prefix prefill, the first launches of a
#      full 32k-token prefill step: RMSNorm, fused-QKV GEMM, tcgen05 prefill attention, O-proj
#      GEMM, RMSNorm, gate/up GEMM (SwiGLU), down GEMM, ... (14 launches = two layers)
#   3. --set full of the decode side on the configs[2] shape (512-token documents, B = 512):
#      tokenizer kernels, sampler, decode attention
set -x
mkdir -p gpurun_out
R=${ROUND:-r02}
HEAD="python bench.py --rows 2048 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --kv-pages 16384"
DOCS="python bench.py --workload docs --rows 512 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --kv-pages 32768 --max-slots 512"
NCU="ncu --clock-control none --kernel-name-base demangled"
T=/tmp/ncu_$R; mkdir -p $T
timeout 900 $NCU --set full -k 'regex:gemm2_bf16_tn_kernel|gemm_bf16_tn_kernel|attn_prefill_tc_kernel|rmsnorm_kernel' -s 256 -c 14 \
    -f -o $T/prefill_step $HEAD > gpurun_out/ncu_prefill_step_${R}.log 2>&1
ncu -i $T/prefill_step.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_prefill_step_raw.csv 2>/dev/null
timeout 600 $NCU --set full --import-source on -k 'regex:gemm2_bf16_tn_kernel<.int.1|attn_prefill_tc_kernel' -s 80 -c 2 \
    -f -o gpurun_out/${R}_down_gemm_and_attn $HEAD > gpurun_out/ncu_src_${R}.log 2>&1
timeout 900 $NCU --set full -k 'regex:attn_decode|sample_greedy|bpe_kernel|pretok_kernel|detok|compact_rows|fsm_build' -c 44 \
    -f -o $T/decode_side $DOCS > gpurun_out/ncu_decode_side_${R}.log 2>&1
ncu -i $T/decode_side.ncu-rep --page raw --csv > gpurun_out/${R}_ncu_decode_side_raw.csv 2>/dev/null
timeout 500 $NCU --metrics gpu__time_duration.sum -c 12000 --csv --log-file gpurun_out/launches_${R}.csv \
    python bench.py --rows 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-secondary --kv-pages 16384 > gpurun_out/ncu_launch_${R}.log 2>&1
tail -n 3 gpurun_out/ncu_prefill_step_${R}.log
tail -n 3 gpurun_out/ncu_decode_side_${R}.log
du -sh gpurun_out; ls -la gpurun_out | tail -12
