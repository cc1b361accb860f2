#!/bin/bash
# One call on an N-GPU box (N = $1, default 8): host->device topology probe, the NCCL multi-rank asserts,
# the bench at N ranks (and at 2 ranks when time allows: $2 = "with2").
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/gpus_${N}.csv
timeout 300 python tools/h2d_probe.py $N > gpurun_out/h2d_probe_${N}gpu.json 2> gpurun_out/h2d_probe_${N}gpu.err; echo "probe rc=$?"; head -c 3000 gpurun_out/h2d_probe_${N}gpu.json | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multirank_check.py > gpurun_out/multirank_check_${N}gpu.json 2> gpurun_out/multirank_check_${N}gpu.err; echo "multirank rc=$?"; tail -c 1500 gpurun_out/multirank_check_${N}gpu.json; tail -5 gpurun_out/multirank_check_${N}gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r2_${N}gpu.json 2> gpurun_out/bench_r2_${N}gpu.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2_${N}gpu.err | cut -c1-600
if [ "$2" = "with2" ]; then
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-paths > gpurun_out/bench_r2_2of${N}gpu.json 2> gpurun_out/bench_r2_2of${N}gpu.err; echo "bench2 rc=$?"
fi
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_r2_${N}gpu.json'))
    print("ms_per_step %.2f value %.3e"%(d['ms_per_step'], d['value']))
    e=d['e2e']; print("e2e ms", e['ms_per_step'], e.get('wall_ms_each_call'), "agg h2d", e['pcie_h2d_gbs_aggregate_concurrent'], "frac", e['frac_of_pcie_bound'], "rank path", e['per_rank_nccl_path']['ms_per_step'])
    print("parity", d['parity'])
    x=d['extra']; print(x.get('stress_c5_sharded')); print(x.get('apsp_sharded_c4'))
except Exception as ex:
    print("parse failed", ex)
PY
