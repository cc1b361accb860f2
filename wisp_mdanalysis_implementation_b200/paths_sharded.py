"""(source, sink) pairs sharded over ranks -- SURVEY 8(e), configs[4].

Every pair of the stress configuration is an independent ``get_paths([s], [t], K)`` call
(network_analysis_functions.py:12 called once per pair), so the pairs are dealt round-robin
to the ranks, each rank runs the all-pairs shortest paths once on its own copy of the cost
matrix (``wisp_paths_prepare``: the matrix is replicated, 32 MB at N = 2,000) and answers its
pairs with ``wisp_paths_query``; the variable-length path lists are gathered on rank 0.  There
is no data-path collective -- only the final gather of results.
"""
from __future__ import annotations

import torch.distributed as dist


def pair_shard(n_pairs, rank, world):
    """Indices of the pairs rank ``rank`` answers (round-robin keeps long and short pairs mixed)."""
    return range(rank, n_pairs, world)


def query_pairs(ctx, w, pairs, number_of_paths, group=None, node0_quirk=True, query=None):
    """``[get_paths(w, [s], [t], K) for (s, t) in pairs]`` computed by all ranks of ``group``.

    Returns the list (in the order of ``pairs``) on rank 0 and ``None`` elsewhere.  ``query`` is
    the per-pair callable; by default the CUDA path of ``ctx`` (one APSP, then one enumeration
    per pair).  The host logic is exercised on gloo with a stand-in callable in
    tests/test_sharding_gloo.py.
    """
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    if query is None:
        ctx.paths_prepare(w)

        def query(s, t):
            return ctx.paths_query([s], [t], number_of_paths, node0_quirk=node0_quirk)

    local = [(i, query(int(pairs[i][0]), int(pairs[i][1]))) for i in pair_shard(len(pairs), rank, world)]
    if world == 1:
        parts = [local]
    else:
        parts = [None] * world if rank == 0 else None
        dist.gather_object(local, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0,
                           group=group)
        if rank != 0:
            return None
    out = [None] * len(pairs)
    for part in parts:
        for i, paths in part:
            out[i] = paths
    return out
