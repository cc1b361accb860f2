#!/bin/bash
# final single-GPU evidence session: tests, bench, launch list, ncu --set full of the dominant kernels
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu -x 2>&1 | tail -6 | tee gpurun_out/full_tests.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2_f.json 2> gpurun_out/bench_r2_f.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_f.err
timeout 300 python tools/fw_time.py 2000 10000 2>&1 | tail -2 | tee gpurun_out/fw_time6.json
PASS="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --parity-frames 0 --apsp-sharded-nodes 0"
$PASS > gpurun_out/plain_pass.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2_final.csv $PASS > gpurun_out/ncu_pass.log 2>&1
echo "launch list rc=$?"
$PASS > gpurun_out/plain_pass2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"contact_sum_ws|com_stream|align_shift|align_apply|align_solve|sum32_seq|gram_i8|center_convert|slice_planes" -s 12 -c 12 -f -o gpurun_out/prof_r2_pass $PASS > gpurun_out/ncu_full_pass.log 2>&1
echo "ncu full pass rc=$?"
python tools/fw_one.py 10000 32 1 > gpurun_out/plain_fw.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"fw32_p3|fw_p2|fw_p1|fw_resolve|fw_chase" -s 200 -c 8 -f -o gpurun_out/prof_r2_fw python tools/fw_one.py 10000 32 1 > gpurun_out/ncu_full_fw.log 2>&1
echo "ncu full fw rc=$?"
ls -la gpurun_out/*.ncu-rep
