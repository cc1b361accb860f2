#!/bin/bash
# Final-evidence session: peaks, bench (default), launch list + full ncu capture of the same command.
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 100 ./tools/peaks > gpurun_out/peaks.json 2>&1; cat gpurun_out/peaks.json
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
( time python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --stress-pairs 4 --apsp-nodes 2000"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-paths"
timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gram_dmma|contact_sum|com_stream|align_shift|center_convert|epilogue' -s 6 -c 8 -o gpurun_out/prof_final -f $CMD2 > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"
