#!/bin/bash
# Final-evidence sessions (ONE profiler use per gpurun call):
#   gpu_final.sh bench     peaks, GPU tests, smoke, default bench
#   gpu_final.sh launches  ncu launch list of a short bench run
#   gpu_final.sh full      ncu --set full capture of one pass of the hot path
export PYTHONPATH=.
mkdir -p gpurun_out
case "$1" in
bench)
  timeout 100 ./tools/peaks > gpurun_out/peaks.json 2>&1; cat gpurun_out/peaks.json
  timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -2
  timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -2
  ( time python bench.py --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real
  ;;
launches)
  CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --stress-pairs 4 --apsp-nodes 2000"
  timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
  echo "launch list rc=$?"
  ;;
full)
  CMD2="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-paths"
  timeout 300 $CMD2 > gpurun_out/plain2.log 2>&1 && \
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gram_i8|gram_dmma|contact_sum|com_stream|align_shift|center_convert|epilogue_kernel|slice_planes' -s 8 -c 8 -o gpurun_out/prof_final -f $CMD2 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full rc=$?"
  ;;
*) echo "usage: $0 bench|launches|full"; exit 2;;
esac
