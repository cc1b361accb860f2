#!/bin/bash
# 2-GPU box: smoke() on cuda:0, then the bench at 2 ranks (strong-scaling row between 1 and 8)
mkdir -p gpurun_out
export PYTHONPATH=.
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke_r2.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 5 --warmup 3 --no-paths --apsp-sharded-nodes 0 > gpurun_out/bench_r2_2gpu_final.json 2> gpurun_out/bench_r2_2gpu_final.err; echo "bench2 rc=$?"; tail -2 gpurun_out/bench_r2_2gpu_final.err | cut -c1-300
