#!/bin/bash
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
for v in 2 4; do
  export WISP_K3_VARIANT=$v
  timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-paths > gpurun_out/bq_$v.json 2> gpurun_out/bq_$v.err
  python -c "
import json; d=json.load(open('gpurun_out/bq_$v.json')); print('variant $v ms_per_step %.2f'%d['ms_per_step'], {k:round(x['ms'],3) for k,x in d['kernels'].items()})"
done
export WISP_K3_VARIANT=4
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-paths --no-e2e --frames 12000"
timeout 300 $CMD > gpurun_out/plain_k3.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'contact_sum' -s 1 -c 1 -o gpurun_out/prof_k3v4 -f $CMD > gpurun_out/ncu_k3.log 2>&1
echo "ncu rc=$?"
