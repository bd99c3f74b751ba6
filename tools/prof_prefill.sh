mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:attn_prefill' -s 60 -c 1 -f -o gpurun_out/prof_attn_prefill_real python bench.py --rows 1024 --steps 1 --warmup 1 --no-cpu-baseline --kv-pages 16384 --max-slots 1024 > gpurun_out/ncu_prefill_real.log 2>&1
tail -2 gpurun_out/ncu_prefill_real.log | cut -c1-200
