"""(source, sink) pairs sharded over ranks -- SURVEY 8(e), configs[4].

Every pair of the stress configuration is an independent ``get_paths([s], [t], K)`` call
(network_analysis_functions.py:12 called once per pair).  The pairs are dealt round-robin to the
ranks; each rank runs the all-pairs shortest paths once on its own copy of the cost matrix
(``wisp_paths_prepare``: 5 ms at N = 2,000, cheaper than shipping the 48 MB result) and answers ALL of
its pairs with ONE batched call (``wisp_paths_query_pairs``: the cutoff passes of every unfinished
query share the same kernels, so the GPU sees ~10 launches per batch instead of ~50 per pair).  The
results travel to rank 0 as four packed arrays per rank (query offsets, lengths, node offsets, node
indices) -- no per-path Python objects are pickled -- and are unpacked lazily.  There is no data-path
collective, only the final gather.
"""
from __future__ import annotations

import numpy as np
import torch.distributed as dist

from . import _lib


def pair_shard(n_pairs, rank, world):
    """Indices of the pairs rank ``rank`` answers (round-robin keeps long and short pairs mixed)."""
    return range(rank, n_pairs, world)


class PackedPathResults:
    """Results of many queries as packed arrays; ``results[i]`` materialises the list of
    ``[length, n0, n1, ...]`` of query i on demand, ``to_lists()`` all of them."""

    def __init__(self, n_queries):
        self._where = [None] * n_queries          # query -> (part index, local query index)
        self._parts = []

    def add_part(self, query_ids, packed):
        k = len(self._parts)
        self._parts.append(packed)
        for local, q in enumerate(query_ids):
            self._where[q] = (k, local)

    def __len__(self):
        return len(self._where)

    def count(self, i):
        k, local = self._where[i]
        qoff = self._parts[k][0]
        return int(qoff[local + 1] - qoff[local])

    def total_paths(self):
        return int(sum(len(p[1]) for p in self._parts))

    def __getitem__(self, i):
        k, local = self._where[i]
        qoff, lengths, noff, nodes = self._parts[k]
        a, b = int(qoff[local]), int(qoff[local + 1])
        ll, no = lengths[a:b].tolist(), noff[a:b + 1].tolist()
        nl = nodes[no[0]:no[-1]].tolist()
        base = no[0]
        return [[ll[j]] + nl[no[j] - base:no[j + 1] - base] for j in range(b - a)]

    def to_lists(self):
        return [self[i] for i in range(len(self))]


def query_pairs(ctx, w, pairs, number_of_paths, group=None, node0_quirk=True, query=None, packed=False):
    """``[get_paths(w, [s], [t], K) for (s, t) in pairs]`` computed by all ranks of ``group``.

    Returns the results (in the order of ``pairs``) on rank 0 and ``None`` elsewhere: a list of lists,
    or with ``packed=True`` a :class:`PackedPathResults`.  ``query`` (tests only) replaces the CUDA
    path with a per-pair callable returning Python lists; the host logic is exercised on gloo with such
    a stand-in in tests/test_sharding_gloo.py.
    """
    distributed = dist.is_available() and dist.is_initialized()
    rank = dist.get_rank(group) if distributed else 0
    world = dist.get_world_size(group) if distributed else 1
    mine = list(pair_shard(len(pairs), rank, world))
    if query is None:
        ctx.paths_prepare(w)
        if mine:
            part = ctx.paths_query_pairs([(int(pairs[i][0]), int(pairs[i][1])) for i in mine], number_of_paths,
                                         node0_quirk=node0_quirk, packed=True)
        else:
            part = (np.zeros(1, np.int64), np.zeros(0), np.zeros(1, np.int64), np.zeros(0, np.int32))
    else:
        part = pack_lists([query(int(pairs[i][0]), int(pairs[i][1])) for i in mine])
    if world == 1:
        parts = [(mine, part)]
    else:
        parts = [None] * world if rank == 0 else None
        dist.gather_object((mine, part), parts, dst=dist.get_global_rank(group, 0) if group is not None else 0,
                           group=group)
        if rank != 0:
            return None
    out = PackedPathResults(len(pairs))
    for ids, packed_part in parts:
        out.add_part(ids, packed_part)
    return out if packed else out.to_lists()


def pack_lists(results):
    """list (per query) of lists of [length, n0, ...] -> (query_offsets, lengths, node_offsets, nodes)."""
    qoff, lengths, noff, nodes = [0], [], [0], []
    for paths in results:
        for p in paths:
            lengths.append(float(p[0]))
            nodes.extend(int(x) for x in p[1:])
            noff.append(len(nodes))
        qoff.append(len(lengths))
    return (np.asarray(qoff, np.int64), np.asarray(lengths, np.float64), np.asarray(noff, np.int64),
            np.asarray(nodes, np.int32))
