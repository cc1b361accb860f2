#!/bin/bash
# One call on an 8-GPU box: the NCCL multi-rank asserts, then the bench at 8 ranks.
N=${1:-8}
mkdir -p gpurun_out
export PYTHONPATH=.
nvidia-smi --query-gpu=index,name,pci.bus_id --format=csv > gpurun_out/gpus_${N}.csv
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multirank_check.py > gpurun_out/multirank_check_${N}gpu.json 2> gpurun_out/multirank_check_${N}gpu.err; echo "multirank rc=$?"; tail -c 1200 gpurun_out/multirank_check_${N}gpu.json; tail -3 gpurun_out/multirank_check_${N}gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r2_${N}gpu_final.json 2> gpurun_out/bench_r2_${N}gpu_final.err; echo "bench rc=$?"; tail -5 gpurun_out/bench_r2_${N}gpu_final.err | cut -c1-600
