#!/bin/bash
# round 2, first call: min-plus issue-rate microbenchmarks + the existing GPU parity suite
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
timeout 200 ./tools/peaks_minplus > gpurun_out/peaks_minplus.json 2>&1; cat gpurun_out/peaks_minplus.json
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -8
