#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multi_shard or query_pairs or duplicate_pairs" 2>&1 | tail -25 | tee gpurun_out/multi_tests.log
timeout 1500 python -m pytest tests -q -m gpu -x -k "not multi_shard and not published and not c3_size" 2>&1 | tail -8 | tee gpurun_out/all_tests.log
timeout 600 python tools/fw_time.py 2000 10000 2>&1 | tail -3 | tee gpurun_out/fw_time2.json
