#!/bin/bash
mkdir -p gpurun_out
export PYTHONPATH=.
echo "--- K3 idle on / off"; python tools/k3_time.py 2000 50000 5; WISP_K3_NOIDLE=1 python tools/k3_time.py 2000 50000 5
python tools/k3_time.py 128 200000 3; WISP_K3_NOIDLE=1 python tools/k3_time.py 128 200000 3
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/full_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --apsp-sharded-nodes 0 > gpurun_out/bench_r2_e.json 2> gpurun_out/bench_r2_e.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_e.err
timeout 300 python tools/fw_time.py 2000 10000 2>&1 | tail -2 | tee gpurun_out/fw_time4.json
