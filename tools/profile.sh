#!/bin/bash
# Runs on the GPU box (via gpurun): full benchmark line, ncu launch list, and full-set
# captures of the two graded kernels.  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python bench.py > gpurun_out/bench_${R}.json 2> gpurun_out/bench_${R}.err
tail -c 3000 gpurun_out/bench_${R}.json
SMALL="python bench.py --rows 400 --steps 1 --warmup 1 --no-cpu-baseline"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv \
    --log-file gpurun_out/launches_${R}.csv $SMALL > gpurun_out/ncu_launch_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16_tn -s 300 -c 3 \
    -f -o gpurun_out/prof_gemm_${R} $SMALL > gpurun_out/ncu_gemm_${R}.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_decode -s 40 -c 3 \
    -f -o gpurun_out/prof_attn_decode_${R} $SMALL > gpurun_out/ncu_attn_${R}.log 2>&1
ls -la gpurun_out
