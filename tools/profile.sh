#!/bin/bash
# Runs on the GPU box (via gpurun).  Stage selectable: bench | launches | gemm | attn | prefill | all
# Outputs land in gpurun_out/.  ncu passes use a small frame and a small KV pool so that
# kernel replay (save/restore of device memory) stays cheap; -s skips the warm-up launches.
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
STAGE=${1:-all}
SMALL="python bench.py --rows 512 --steps 1 --warmup 1 --no-cpu-baseline --kv-pages 8192 --max-slots 512"
NCU="ncu --clock-control none --kernel-name-base demangled"
if [[ $STAGE == bench || $STAGE == all ]]; then
  timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
  tail -c 3500 gpurun_out/bench_${R}.json
fi
if [[ $STAGE == launches || $STAGE == all ]]; then
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 16000 --csv \
      --log-file gpurun_out/launches_${R}.csv $SMALL > gpurun_out/ncu_launch_${R}.log 2>&1
fi
if [[ $STAGE == gemm || $STAGE == all ]]; then
  # prefill-shape launches of the CTA-pair GEMM: SwiGLU (gate/up) and fused QKV epilogues
  timeout 420 $NCU --set full --import-source on -k 'regex:gemm2_bf16_tn_kernel<.int.2' -s 4 -c 2 \
      -f -o gpurun_out/prof_gemm_gateup_${R} $SMALL > gpurun_out/ncu_gemm_${R}.log 2>&1
  timeout 420 $NCU --set full --import-source on -k 'regex:gemm2_bf16_tn_kernel<.int.4' -s 4 -c 2 \
      -f -o gpurun_out/prof_gemm_qkv_${R} $SMALL >> gpurun_out/ncu_gemm_${R}.log 2>&1
fi
if [[ $STAGE == attn || $STAGE == all ]]; then
  timeout 420 $NCU --set full --import-source on -k 'regex:attn_decode' -s 40 -c 3 \
      -f -o gpurun_out/prof_attn_decode_${R} $SMALL > gpurun_out/ncu_attn_${R}.log 2>&1
fi
if [[ $STAGE == prefill || $STAGE == all ]]; then
  timeout 420 $NCU --set full --import-source on -k 'regex:attn_prefill' -s 4 -c 2 \
      -f -o gpurun_out/prof_attn_prefill_${R} $SMALL > gpurun_out/ncu_prefill_${R}.log 2>&1
fi
ls -la gpurun_out
