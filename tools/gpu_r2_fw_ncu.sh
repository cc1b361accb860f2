#!/bin/bash
mkdir -p gpurun_out
python tools/fw_one.py 10000 32 1 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 460 --csv --log-file gpurun_out/fw32_10000_launches.csv python tools/fw_one.py 10000 32 1 > gpurun_out/ncu.log 2>&1
python tools/fw_one.py 2000 64 1 > gpurun_out/plain2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/fw64_2000_launches.csv python tools/fw_one.py 2000 64 1 > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu.log gpurun_out/ncu2.log
