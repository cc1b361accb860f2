#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "pipeline_vs_oracle or tiny or known_answers or golden_trajectory or properties_large or published_config_c2_vs" 2>&1 | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-paths --apsp-sharded-nodes 0 > gpurun_out/bench_r2_d.json 2> gpurun_out/bench_r2_d.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_d.err
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_pass.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0 > gpurun_out/ncu_pass.log 2>&1
echo "ncu rc=$?"
