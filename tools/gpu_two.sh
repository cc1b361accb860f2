#!/bin/bash
# 2-GPU session: GPU tests, then the bench with the path stage (pairs sharded over both ranks).
export PYTHONPATH=.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 \
    bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline --apsp-nodes 0 > gpurun_out/bench_2gpu_paths.json 2> gpurun_out/bench_2gpu_paths.err
echo "rc=$?"
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench_2gpu_paths.json'))
    print("ms_per_step %.3f e2e %.1f"%(d['ms_per_step'], d['e2e']['ms_per_step']))
    print({k:v for k,v in d['extra'].items() if 'stress' in k})
except Exception as e:
    print("parse failed", e); print(open('gpurun_out/bench_2gpu_paths.err').read()[-2500:])
PY
