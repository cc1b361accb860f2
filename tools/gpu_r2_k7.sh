#!/bin/bash
# round 2, second session: the full GPU suite behind per-test timeouts, bench, launch list
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 700 python -u tools/gpu_tests_resilient.py --per-test 150 --budget 600 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-paths --apsp-sharded-nodes 0 > gpurun_out/bench_r2_k.json 2> gpurun_out/bench_r2_k.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_k.err
PASS="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0"
timeout 300 $PASS > gpurun_out/plain_pass.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_k7.csv $PASS > gpurun_out/ncu_pass.log 2>&1
echo "launch list rc=$?"
