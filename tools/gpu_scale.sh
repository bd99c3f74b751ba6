#!/usr/bin/env bash
# 8 GPUs: the sharded frame path under torchrun exactly as the driver's SCALE run launches it.
set -u
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 \
    bench.py --gpus $N --steps 2 --warmup 1 > gpurun_out/f2_bench_n$N.json 2> gpurun_out/f2_bench_n$N.err
echo "bench N=$N rc=$?"; tail -6 gpurun_out/f2_bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/f2_bench_n{n}.json"))
    print({k: d[k] for k in ("value", "ms_per_step", "n_gpus", "scaling", "gpu_launches")}, "e2e", d["e2e"]["value"])
    print(d.get("sharding")); print("valid", d["outputs_valid"], d["clocks"])
except Exception as e:
    print("parse failed", e)
PY
