#!/bin/bash
# round 2, second session: the full GPU suite behind per-test timeouts, K3 variants, bench
mkdir -p gpurun_out
timeout 1300 python -u tools/gpu_tests_resilient.py --per-test 200 --budget 1200 2>&1 | tail -3
for v in 58 68; do WISP_K3_VARIANT=$v timeout 120 python tools/k3_time.py 2000 50000 3 2>&1 | tail -1 | tee -a gpurun_out/k3_variants.log; done
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-paths --apsp-sharded-nodes 0 > gpurun_out/bench_r2_g.json 2> gpurun_out/bench_r2_g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_g.err
