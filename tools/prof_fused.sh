mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:gemm2_bf16_tn_kernel<.int.4' -s 40 -c 1 -f -o gpurun_out/prof_gemm_qkv_fused python bench.py --rows 256 --steps 1 --warmup 1 --no-cpu-baseline --kv-pages 8192 --max-slots 512 > gpurun_out/ncu_qkv_fused.log 2>&1
tail -3 gpurun_out/ncu_qkv_fused.log | cut -c1-300
