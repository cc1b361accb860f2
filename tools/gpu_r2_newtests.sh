#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "published_config or md_like or c3_size" 2>&1 | tail -25 | tee gpurun_out/newtests.log
