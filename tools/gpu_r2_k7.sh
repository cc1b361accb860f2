#!/bin/bash
# round 2, second session: the full GPU suite behind per-test timeouts, K3 variants, bench, launch list
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 1000 python -u tools/gpu_tests_resilient.py --per-test 150 --budget 900 2>&1 | tail -3
for v in 58 68; do WISP_K3_VARIANT=$v timeout 120 python tools/k3_time.py 2000 50000 3 2>&1 | tail -1 | tee -a gpurun_out/k3_variants.log; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-paths --apsp-sharded-nodes 0 > gpurun_out/bench_r2_h.json 2> gpurun_out/bench_r2_h.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_h.err
PASS="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0"
timeout 300 $PASS > gpurun_out/plain_pass.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_k7.csv $PASS > gpurun_out/ncu_pass.log 2>&1
echo "launch list rc=$?"
