#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/full_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_b.json 2> gpurun_out/bench_r2_b.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_b.err
timeout 300 python tools/fw_time.py 2000 10000 2>&1 | tail -2 | tee gpurun_out/fw_time3.json
