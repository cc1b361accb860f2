"""torchrun --nproc-per-node R tools/apsp_sharded_bench.py [N]: row-block sharded FP32 APSP
(configs[3]).  Rank 0 prints one JSON line; with N <= 4096 the result is checked against the
single-GPU kernel."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisp_mdanalysis_implementation_b200 import _lib  # noqa: E402
from wisp_mdanalysis_implementation_b200.apsp_sharded import ShardedAPSP, TILE  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dev = torch.device("cuda", local_rank)
ctx = _lib.default_context(local_rank)
sh = ShardedAPSP(ctx, N, rank=rank, world=world)
r0, r1 = sh.t0 * TILE, min(N, sh.t1 * TILE)
# symmetric random costs, generated row-block-wise from a counter-based hash so every rank agrees
def rows(a, b):
    i = torch.arange(a, b, device=dev, dtype=torch.int64)[:, None]
    j = torch.arange(0, N, device=dev, dtype=torch.int64)[None, :]
    lo, hi = torch.minimum(i, j), torch.maximum(i, j)
    h = (lo * 1000003 + hi * 7919 + 12345) % 2147483647
    h = (h * 48271) % 2147483647
    w = 0.05 + 2.0 * (h.to(torch.float64) / 2147483647.0)
    w[i == j] = 0.0
    return w
d_rows = rows(r0, r1)
times = []
for rep in range(2):
    sh.load_rows(d_rows)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    sh.run()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ok = None
if N <= 4096:
    full = rows(0, N).cpu().numpy()
    want, _ = ctx.apsp(full, precision=32, want_pred=False)
    ok = bool(np.array_equal(sh.local_rows().cpu().numpy().astype(np.float64), want[r0:r1]))
    if world > 1:
        f = torch.tensor([1.0 if ok else 0.0], device=dev)
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        ok = bool(f.item() > 0.5)
if rank == 0:
    ms = min(times)
    print(json.dumps({"metric": "row-block sharded FP32 Floyd-Warshall wall time", "nodes": N, "n_gpus": world,
                      "ms": ms, "trelax_per_s": N ** 3 / (ms * 1e-3) / 1e12, "rounds": sh.n_tiles,
                      "panel_bytes": TILE * sh.n_pad * 4, "matches_single_gpu": ok}))
if world > 1:
    dist.destroy_process_group()
