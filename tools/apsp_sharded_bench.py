"""torchrun --nproc-per-node R tools/apsp_sharded_bench.py [N]: row-block sharded FP32 APSP (configs[3]).
Rank 0 prints one JSON line (timing with / without look-ahead + the bit-equality check against the single-GPU
matrix at 8,192 nodes on the same ranks)."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisp_mdanalysis_implementation_b200 import _lib  # noqa: E402
from wisp_mdanalysis_implementation_b200.apsp_sharded import run_sharded_benchmark  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
res = run_sharded_benchmark(_lib.default_context(local_rank), N, rank, world)
if rank == 0:
    res["metric"] = "row-block sharded FP32 Floyd-Warshall wall time"
    print(json.dumps(res))
if world > 1:
    dist.destroy_process_group()
