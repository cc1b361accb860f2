#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/relay_check.py 2 2>&1 | tail -3 | tee gpurun_out/relay_check_2gpu.json
timeout 600 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "multirank_nccl or multi_shard_with" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --apsp-sharded-nodes 16384 --apsp-nodes 0 > gpurun_out/bench_r2_2gpu.json 2> gpurun_out/bench_r2_2gpu.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r2_2gpu.err | cut -c1-400
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r2_2gpu.json'))
print("ms_per_step %.2f value %.3e"%(d['ms_per_step'], d['value']))
for k,v in d['kernels'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a not in ('note','residuals')})
e=d['e2e']; print("e2e ms", e['ms_per_step'], [round(x) for x in e['wall_ms_each_call']], e['stages_of_each_call_ms'][0], "rank path", e['per_rank_nccl_path']['ms_per_step'])
print("parity", d['parity'])
x=d['extra']; print(x.get('north_star')); print(x.get('stress_c5_sharded')); print(x.get('apsp_sharded_c4'))
PY
