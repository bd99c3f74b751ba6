#!/usr/bin/env bash
# 2 GPUs: the sharded frame path under torchrun (what the driver's SCALE run does), then the
# in-process multi-GPU test
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus 2 --steps 2 --warmup 2 > gpurun_out/c5_bench_n2.json 2> gpurun_out/c5_bench_n2.err
echo "bench N=2 rc=$?"; tail -6 gpurun_out/c5_bench_n2.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/c5_bench_n2.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "gpu_launches")}, "e2e", d["e2e"]["value"])
    print(d["sharding"]); print(d["e2e"]); print("valid", d["outputs_valid"])
except Exception as e:
    print("parse failed", e)
PY
timeout 600 python -m pytest tests/test_engine_gpu.py -q -k "multi_gpu" --timeout 600 -p no:cacheprovider 2>&1 | tail -4
