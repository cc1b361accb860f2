#!/bin/bash
# K3 with the active blocks of diagonal / padded tiles dealt out evenly over the schedulers: full GPU suite
# on the new default, then K3 alone with and without the balancing (same build, WISP_K3_BALANCE)
mkdir -p gpurun_out
export PYTHONPATH=.
for b in 0 1; do WISP_K3_BALANCE=$b timeout 100 python tools/k3_time.py 2000 50000 5 2>&1 | tail -1 | sed "s/^/balance=$b /" | tee -a gpurun_out/k3_balance.log; done
timeout 330 python -u tools/gpu_tests_resilient.py --per-test 100 --budget 300 2>&1 | tail -3
