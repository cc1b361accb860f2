"""torchrun --nproc-per-node R tools/multirank_check.py: asserts the multi-rank paths on real NCCL ranks
(ADVICE r1): (1) ShardedAPSP rows == single-GPU wisp_apsp rows, with and without look-ahead;
(2) FramePipeline.run_device / run_host (K7 alignment on) and run_host_streamed (alignment off),
all-reduced over the ranks, == the single-GPU wisp_pipeline; (3) paths_sharded.query_pairs == serial
get_paths; (4) on rank 0, wisp_pipeline on an in-process context over all R GPUs == single GPU.
Rank 0 prints one JSON line; any mismatch raises."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisp_mdanalysis_implementation_b200 import _lib, paths_sharded, synth  # noqa: E402
from wisp_mdanalysis_implementation_b200.apsp_sharded import ShardedAPSP, TILE  # noqa: E402
from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local_rank = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local_rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
dev = torch.device("cuda", local_rank)
ctx = _lib.default_context(local_rank)
report = {"ranks": world}


def all_ok(flag):
    t = torch.tensor([1.0 if flag else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item() > 0.5)


# ---- (1) sharded APSP ----
n = 3000
W = synth.random_cost_matrix(n, seed=44, density=0.01)
want, _ = ctx.apsp(W, precision=32, want_pred=False)
d_W = torch.from_numpy(W).to(dev)
for la in (True, False):
    sh = ShardedAPSP(ctx, n, rank=rank, world=world)
    sh.load_rows(d_W[sh.t0 * TILE:min(n, sh.t1 * TILE)].contiguous())
    sh.run(lookahead=la)
    torch.cuda.synchronize()
    got = sh.local_rows().cpu().numpy().astype(np.float64)
    ok = all_ok(np.array_equal(got, want[sh.t0 * TILE:min(n, sh.t1 * TILE)]))
    report["sharded_apsp_bit_equal_lookahead_%s" % la] = ok
    assert ok, "sharded APSP differs from the single-GPU matrix (lookahead=%s)" % la

# ---- (2) frame-sharded pipelines ----
N, T, apn, cutoff = 300, 2400, 5, 13.0
system, coords = synth.synth(N, T, apn, seed=12)
ang = np.linspace(0.0, 0.5, T)
c, s = np.cos(ang), np.sin(ang)
rot = np.zeros((T, 3, 3)); rot[:, 0, 0] = c; rot[:, 0, 1] = -s; rot[:, 1, 0] = s; rot[:, 1, 1] = c; rot[:, 2, 2] = 1
centre = coords.mean(axis=1, keepdims=True)
coords = (np.einsum("tac,tdc->tad", coords - centre, rot) + centre).astype(np.float32)
t0, t1 = T * rank // world, T * (rank + 1) // world
for aligned in (True, False):
    ref = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=cutoff,
                       which_map=1, maximum_num_iterations=100 if aligned else 0)
    pipe = FramePipeline(ctx, N, system.n_atoms, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                         frames_local=t1 - t0, frames_total=T, cutoff=cutoff, which_map=1,
                         align_iterations=100 if aligned else 0)
    h = torch.from_numpy(coords[t0:t1]).pin_memory()
    modes = [("run_device", lambda: pipe.run_device(h.to(dev))), ("run_host", lambda: pipe.run_host(h, fetch=False))]
    if not aligned:
        modes.append(("run_host_streamed", lambda: pipe.run_host_streamed(h, n_blocks=3, fetch=False)))
    for name, fn in modes:
        fn()
        torch.cuda.synchronize()
        corr, avgdist, binary = pipe.d_corr.cpu().numpy(), pipe.d_avgdist.cpu().numpy(), pipe.d_binary.cpu().numpy()
        e_c = float((np.abs(corr - ref["corr"]) / np.maximum(np.abs(ref["corr"]), 1e-3)).max())
        e_d = float((np.abs(avgdist - ref["avgdist"]) / np.maximum(ref["avgdist"], 1e-9))[~np.eye(N, dtype=bool)].max())
        near = np.abs(ref["avgdist"] - cutoff) <= 1e-6 * cutoff      # ties with the cutoff may fall either way
        mism = int(((binary != ref["binary"]) & ~near).sum())
        ok = all_ok(e_c <= 1e-6 and e_d <= 1e-6 and mism == 0)
        report["%s_aligned_%s" % (name, aligned)] = {"max_rel_corr": e_c, "max_rel_dist": e_d, "topology_mismatches": mism}
        assert ok, (name, aligned, e_c, e_d, mism)
        if aligned:
            assert len(pipe.alignment_log) == len(ref["alignment_log"])

# ---- (3) pair-sharded path queries ----
Wp = synth.random_cost_matrix(60, seed=12, density=0.15, lo=0.05, hi=0.9)
pairs = [(3, 41), (0, 17), (29, 5), (11, 48), (7, 52), (33, 2), (19, 44)]
got = paths_sharded.query_pairs(ctx, Wp, pairs, 9)
if rank == 0:
    for (a, b), paths in zip(pairs, got):
        assert paths == ctx.get_paths(Wp, [a], [b], 9), (a, b)
    report["query_pairs_equal_serial"] = True
dist.barrier()

# ---- (4) in-process multi-GPU context on rank 0 ----
if rank == 0:
    mctx = _lib.Context(gpus=world)
    ref = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=cutoff,
                       which_map=1, maximum_num_iterations=100)
    out = mctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=cutoff,
                        which_map=1, maximum_num_iterations=100)
    e_c = float((np.abs(out["corr"] - ref["corr"]) / np.maximum(np.abs(ref["corr"]), 1e-3)).max())
    e_d = float((np.abs(out["avgdist"] - ref["avgdist"]) / np.maximum(ref["avgdist"], 1e-9))[~np.eye(N, dtype=bool)].max())
    near = np.abs(ref["avgdist"] - cutoff) <= 1e-6 * cutoff
    mism = int(((out["binary"] != ref["binary"]) & ~near).sum())
    report["inprocess_multi_gpu_pipeline"] = {"gpus": mctx.n_gpus, "max_rel_corr": e_c, "max_rel_dist": e_d,
                                              "topology_mismatches": mism}
    assert e_c <= 1e-6 and e_d <= 1e-6 and mism == 0
    mctx.close()
    print(json.dumps(report))
dist.barrier()
dist.destroy_process_group()
