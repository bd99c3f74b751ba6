#!/usr/bin/env bash
# round 2, GPU call 1: the new tcgen05 prefill attention (parity + speed), then the whole suite
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -1
python -c "import os; print('cpus', os.cpu_count())"; free -g | head -2
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -k "dense" 2>&1 | tail -15
timeout 120 python tools/attn_bench.py 2>&1 | tail -6
timeout 120 python tools/attn_bench.py --new 150 --past 46 2>&1 | tail -4
timeout 120 python tools/attn_bench.py --new 512 --past 0 2>&1 | tail -4
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -x -q 2>&1 | tail -15
