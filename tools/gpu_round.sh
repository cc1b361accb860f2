#!/bin/bash
# One GPU session: parity tests, smoke, bench, ncu launch list + full capture of the hot kernels.
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
timeout 120 python __graft_entry__.py --smoke 2>&1 | tail -3
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-paths --no-e2e"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
timeout 300 $CMD > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gram_dmma|contact_sum|com_kernel' -s 3 -c 3 -o gpurun_out/prof_k123 -f $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out | head -20
