#!/bin/bash
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for m in two-pass streamed; do
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-paths --e2e-mode $m > gpurun_out/bench_e2e_$m.json 2> gpurun_out/bench_e2e_$m.err; echo "rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_e2e_$m.json')); print('$m', 'step %.2f ms'%d['ms_per_step'], 'e2e %.1f ms'%d['e2e']['ms_per_step'], d['e2e']['api'][:40])"
done
tail -3 gpurun_out/bench_e2e_streamed.err
