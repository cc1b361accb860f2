#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "apsp or get_paths or smoke or paths" 2>&1 | tail -15 | tee gpurun_out/fw_tests.log
timeout 600 python tools/fw_time.py 2000 10000 2>&1 | tail -3 | tee gpurun_out/fw_time.json
