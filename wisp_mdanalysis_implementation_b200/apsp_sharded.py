"""Row-block sharded FP32 Floyd-Warshall over the GPUs of one box (BASELINE configs[3]:
30,000 nodes, APSP row-block sharded across 8 GPUs with an NCCL broadcast of the pivot block
per round).

Rank r owns the contiguous block of 128-row tile rows ``tile_range(r)`` of the distance matrix.
Round kb (pivot = tile row kb):
  owner(kb):  copy its pivot rows into the panel buffer, finish the panel (phase 1 + pivot-row
              tiles of phase 2), broadcast it -- 128 x Npad floats, 15 MB at N = 30,000;
  every rank: pivot-column tiles + remainder tiles of its own rows from the received panel
              (the owner skips its pivot rows and copies the finished panel back into them).
The kernels are the single-GPU FP32 kernels of csrc/k5_apsp.cu with the row panel passed as a
separate pointer.  ``ShardedAPSP.round_pivot`` / ``round_update`` are exposed separately so the
protocol can be driven rank by rank inside one process (tests) as well as over NCCL (``run``).
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib

TILE = 128


class ShardedAPSP:
    def __init__(self, ctx, n_nodes, rank=0, world=1, group=None):
        self.ctx, self.N, self.rank, self.world, self.group = ctx, int(n_nodes), int(rank), int(world), group
        self.dev = torch.device("cuda", ctx.device)
        self.n_pad = _lib.gram_ld(self.N)
        self.n_tiles = self.n_pad // TILE
        self.t0, self.t1 = self.tile_range(self.rank)
        self.rows_pad = (self.t1 - self.t0) * TILE
        self.d_dist = torch.empty((max(self.rows_pad, TILE), self.n_pad), dtype=torch.float32, device=self.dev)
        self.d_panel = torch.empty((TILE, self.n_pad), dtype=torch.float32, device=self.dev)

    def tile_range(self, rank):
        return self.n_tiles * rank // self.world, self.n_tiles * (rank + 1) // self.world

    def owner(self, kb):
        for r in range(self.world):
            a, b = self.tile_range(r)
            if a <= kb < b:
                return r
        raise ValueError(kb)

    def _stream(self):
        return torch.cuda.current_stream(self.dev).cuda_stream or 1

    def load_rows(self, d_w_rows):
        """d_w_rows: float64 device tensor holding rows [t0*128, min(N, t1*128)) of the cost matrix."""
        if self.rows_pad:
            self.ctx.dev_fw32_init_rows(d_w_rows.data_ptr(), self.N, self.t0 * TILE, self.rows_pad,
                                        self.d_dist.data_ptr(), stream=self._stream())

    def round_pivot(self, kb):
        """Owner only: finish the pivot row panel of round kb in self.d_panel."""
        lt = kb - self.t0
        self.d_panel.copy_(self.d_dist[lt * TILE:(lt + 1) * TILE])
        self.ctx.dev_fw32_pivot_panel(self.d_panel.data_ptr(), self.N, kb, stream=self._stream())

    def round_update(self, kb, panel=None):
        """Every rank: apply the (received) panel of round kb to the local rows."""
        panel = self.d_panel if panel is None else panel
        if not self.rows_pad:
            return
        lt = kb - self.t0 if self.t0 <= kb < self.t1 else -1
        self.ctx.dev_fw32_update_rows(self.d_dist.data_ptr(), self.rows_pad, panel.data_ptr(), self.N, kb,
                                      lt, stream=self._stream())
        if lt >= 0:
            self.d_dist[lt * TILE:(lt + 1) * TILE].copy_(panel)

    def run(self):
        """All rounds over torch.distributed (NCCL): one broadcast of the pivot panel per round."""
        for kb in range(self.n_tiles):
            src = self.owner(kb)
            if src == self.rank:
                self.round_pivot(kb)
            if self.world > 1:
                dist.broadcast(self.d_panel, src=src, group=self.group)
            self.round_update(kb)

    def local_rows(self):
        """float32 distances of the rows this rank owns (padding rows / columns stripped)."""
        r0 = self.t0 * TILE
        r1 = min(self.N, self.t1 * TILE)
        return self.d_dist[:max(0, r1 - r0), :self.N]
