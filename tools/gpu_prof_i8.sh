#!/bin/bash
# ncu capture of the INT8 K2 stage (statistics, digit planes, tcgen05 Gram) at C2 size.
export PYTHONPATH=.
mkdir -p gpurun_out
CMD="python tools/i8gram_test.py 2000 50000 2"
timeout 200 $CMD > gpurun_out/plain_i8.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gram_i8|slice_planes|node_stats|node_scale' -c 4 -o gpurun_out/prof_i8 -f $CMD > gpurun_out/ncu_i8.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_i8.log
