#!/bin/bash
# Everything the driver runs at round end, on one box: GPU tests, smoke, default bench.
mkdir -p gpurun_out
R=${ROUND:-r01}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_${R}_final.json 2> gpurun_out/bench_${R}_final.err
tail -12 gpurun_out/bench_${R}_final.err
python -c "
import json; d=json.load(open('gpurun_out/bench_${R}_final.json'))
print('value',d['value'],'e2e',d['e2e']['value'],'out tok/s',d['output_tokens_per_sec'])
print('roofline',d['roofline']['achieved'],d['roofline']['frac'],'attn',d['roofline_attn_decode']['achieved'],d['roofline_attn_decode']['frac'])
print('cpu',d['cpu_baseline']); print('clocks',d['clocks']); print(d['kernel_ms_profiled_step'])"
