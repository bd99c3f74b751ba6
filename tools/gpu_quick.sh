#!/usr/bin/env bash
# two-minute sanity: kernel + tokenizer + public-API GPU tests and smoke()
set -u
timeout 400 python -m pytest tests/test_kernels_gpu.py tests/test_public_api_gpu.py tests/test_c_host_gpu.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m pytest tests/test_engine_gpu.py -q -k "tokenizer" -p no:cacheprovider 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
