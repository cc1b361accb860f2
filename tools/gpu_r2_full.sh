#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 | tee gpurun_out/full_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_r2_c.json 2> gpurun_out/bench_r2_c.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_c.err
