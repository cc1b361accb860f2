#!/bin/bash
# usage: gpu_prof.sh <kernel-regex> <outname> [bench args]
export PYTHONPATH=.
mkdir -p gpurun_out
K="$1"; O="$2"; shift 2
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --frames 12000 $@"
timeout 300 $CMD > gpurun_out/plain_$O.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$K" -s 2 -c 2 -o gpurun_out/prof_$O -f $CMD > gpurun_out/ncu_$O.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$O.log
