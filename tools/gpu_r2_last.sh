#!/bin/bash
# the resident pass of the final build (K3 balancing on), short: value + per-kernel timings only
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 110 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0 > gpurun_out/bench_r2_final_pass.json 2> gpurun_out/bench_r2_final_pass.err; echo "rc=$?"; tail -2 gpurun_out/bench_r2_final_pass.err | cut -c1-300
