#!/bin/bash
# round 2 final single-GPU evidence: the default bench line, launch list, ncu --set full of one pass
# (the GPU suite ran on this build in the previous call: tools/gpu_r2_k7.sh, 64 passed)
mkdir -p gpurun_out
export PYTHONPATH=.
if [ "$1" = "tests" ]; then timeout 700 python -u tools/gpu_tests_resilient.py --per-test 150 --budget 600 2>&1 | tail -3; fi
timeout 900 python bench.py > gpurun_out/bench_r2_final_1gpu.json 2> gpurun_out/bench_r2_final_1gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_final_1gpu.err
PASS="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0"
timeout 300 $PASS > gpurun_out/plain_pass.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_final.csv $PASS > gpurun_out/ncu_pass.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"contact_sum_ws|com_stream|align_shift|align_apply|align_corr|align_quat|sum32_tma|gram_i8|center_convert|slice_planes" -c 30 -f -o gpurun_out/prof_r2_final $PASS > gpurun_out/ncu_full_pass.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_r2_final.ncu-rep
