#!/bin/bash
# K3 with the active blocks of diagonal / padded tiles dealt out evenly over the schedulers: K3 alone with and
# without the balancing (same build, WISP_K3_BALANCE), then the K3-relevant part of the GPU suite
mkdir -p gpurun_out
export PYTHONPATH=.
for b in 0 1; do WISP_K3_BALANCE=$b timeout 100 python tools/k3_time.py 2000 50000 5 2>&1 | tail -1 | sed "s/^/balance=$b /" | tee -a gpurun_out/k3_balance.log; done
timeout 150 python -m pytest tests -m gpu -q -x -k "pipeline_vs_oracle or published_config_c2_vs_oracle or tiny_and_ragged or properties_large or frame_pipeline_matches or multi_shard_without" 2>&1 | tail -3 | tee gpurun_out/k3_balance_tests.log
