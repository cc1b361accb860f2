#!/bin/bash
# Strong-scaling bench lines at N = 2, 4, 8 GPUs of one box (run with gpurun --gpus 8).
export PYTHONPATH=.
mkdir -p gpurun_out
for n in 2 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) \
      bench.py --gpus $n --steps 5 --warmup 3 --no-cpu-baseline --no-paths > gpurun_out/bench_${n}gpu.json 2> gpurun_out/bench_${n}gpu.err
  echo "N=$n rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${n}gpu.json'))
    print("N=$n ms_per_step %.3f value %.3e e2e %.1f ms"%(d['ms_per_step'], d['value'], d['e2e']['ms_per_step']), {k:round(v['ms'],3) for k,v in d['kernels'].items()})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_${n}gpu.err').read()[-1500:])
PY
done
