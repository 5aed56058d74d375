#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_multiembed_gpu.py tests/test_rot_corr_gpu.py tests/test_prune_gpu.py tests/test_cyclical3_embed_gpu.py tests/test_refining_gpu.py -x -q 2>&1 | tail -15
