"""Row-block sharded FP32 Floyd-Warshall over the GPUs of one box (BASELINE configs[3]:
30,000 nodes, APSP row-block sharded across 8 GPUs with an NCCL broadcast of the pivot block
per round).

Rank r owns the contiguous block of 128-row tile rows ``tile_range(r)`` of the distance matrix.
Round kb (pivot = tile row kb):
  owner(kb):  copy its pivot rows into the panel buffer, finish the panel (phase 1 + pivot-row
              tiles of phase 2), broadcast it -- 128 x Npad floats, 15 MB at N = 30,000;
  every rank: pivot-column tiles + remainder tiles of its own rows from the received panel
              (the owner skips its pivot rows and copies the finished panel back into them).
LOOK-AHEAD (``run``): the owner of the NEXT pivot rows updates that tile-row first, then builds the
next panel and broadcasts it on a second, high-priority stream while every rank is still busy with
the rest of the current round; two panel buffers alternate.  The 15 MB broadcast and the serial
pivot phases leave the critical path.
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
        self.d_panel2 = torch.empty((TILE, self.n_pad), dtype=torch.float32, device=self.dev)
        self.side = torch.cuda.Stream(device=self.dev, priority=-1)

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

    def _build_panel(self, kb, panel):
        lt = kb - self.t0
        panel.copy_(self.d_dist[lt * TILE:(lt + 1) * TILE])
        self.ctx.dev_fw32_pivot_panel(panel.data_ptr(), self.N, kb, stream=self._stream())

    def _update(self, kb, panel, lt, lt_next, part):
        if self.rows_pad:
            self.ctx.dev_fw32_update_rows_part(self.d_dist.data_ptr(), self.rows_pad, panel.data_ptr(), self.N, kb,
                                               lt, lt_next, part, stream=self._stream())

    def run(self, lookahead=True):
        """All rounds over torch.distributed (NCCL): one broadcast of the pivot panel per round, hidden
        behind the previous round's remainder update when ``lookahead``."""
        if self.world == 1 or not lookahead:
            for kb in range(self.n_tiles):
                src = self.owner(kb)
                if src == self.rank:
                    self.round_pivot(kb)
                if self.world > 1:
                    dist.broadcast(self.d_panel, src=src, group=self.group)
                self.round_update(kb)
            return
        main = torch.cuda.current_stream(self.dev)
        side = self.side
        panels = [self.d_panel, self.d_panel2]
        ev_lead, ev_bcast = torch.cuda.Event(), torch.cuda.Event()
        src = self.owner(0)
        if src == self.rank:
            self._build_panel(0, panels[0])
        dist.broadcast(panels[0], src=src, group=self.group)
        for kb in range(self.n_tiles):
            panel = panels[kb & 1]
            lt = kb - self.t0 if self.t0 <= kb < self.t1 else -1
            if lt >= 0:      # the finished pivot rows go back in place before anything reads them again
                self.d_dist[lt * TILE:(lt + 1) * TILE].copy_(panel)
            nxt = kb + 1
            if nxt >= self.n_tiles:
                self._update(kb, panel, lt, -1, 0)
                break
            nsrc = self.owner(nxt)
            mine = nsrc == self.rank
            lt_next = nxt - self.t0 if mine else -1
            if mine:
                self._update(kb, panel, lt, lt_next, 1)          # only the next pivot rows
            ev_lead.record(main)
            with torch.cuda.stream(side):
                side.wait_event(ev_lead)
                if mine:
                    self._build_panel(nxt, panels[nxt & 1])
                dist.broadcast(panels[nxt & 1], src=nsrc, group=self.group)
                ev_bcast.record(side)
            self._update(kb, panel, lt, lt_next, 2 if mine else 0)   # everything else, under the broadcast
            main.wait_event(ev_bcast)

    def local_rows(self):
        """float32 distances of the rows this rank owns (padding rows / columns stripped)."""
        r0 = self.t0 * TILE
        r1 = min(self.N, self.t1 * TILE)
        return self.d_dist[:max(0, r1 - r0), :self.N]


def synthetic_cost_rows(n_nodes, row0, row1, device):
    """Rows [row0, row1) of a dense symmetric random cost matrix, generated from a counter-based hash so
    that every rank produces consistent rows without communication (configs[3] input)."""
    i = torch.arange(row0, row1, device=device, dtype=torch.int64)[:, None]
    j = torch.arange(0, n_nodes, device=device, dtype=torch.int64)[None, :]
    lo, hi = torch.minimum(i, j), torch.maximum(i, j)
    h = (lo * 1000003 + hi * 7919 + 12345) % 2147483647
    h = (h * 48271) % 2147483647
    w = 0.05 + 2.0 * (h.to(torch.float64) / 2147483647.0)
    w[i == j] = 0.0
    return w


def run_sharded_benchmark(ctx, n_nodes, rank, world, group=None, check_nodes=8192, reps=2):
    """configs[3]: time the row-block sharded Floyd-Warshall at ``n_nodes`` with and without look-ahead
    (device time, max over ranks) and CHECK it on the hardware it ran on: the same protocol at
    ``check_nodes`` must equal the single-GPU matrix bit for bit.  Returns a dict on every rank."""
    dev = torch.device("cuda", ctx.device)

    def max_over_ranks(x):
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        return float(t.item())

    out = {"nodes": int(n_nodes), "n_gpus": int(world)}
    # ---- correctness on this hardware ----
    nc = int(min(check_nodes, n_nodes))
    sh = ShardedAPSP(ctx, nc, rank=rank, world=world, group=group)
    r0, r1 = sh.t0 * TILE, min(nc, sh.t1 * TILE)
    sh.load_rows(synthetic_cost_rows(nc, r0, r1, dev))
    sh.run(lookahead=True)
    full = synthetic_cost_rows(nc, 0, nc, dev)
    want = torch.empty((nc, nc), dtype=torch.float64, device=dev)
    ctx.dev_apsp(full.data_ptr(), nc, 32, want.data_ptr(), None, stream=torch.cuda.current_stream(dev).cuda_stream or 1)
    torch.cuda.synchronize()
    same = bool(torch.equal(sh.local_rows().to(torch.float64), want[r0:r1]))
    flag = torch.tensor([1.0 if same else 0.0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
    out["matches_single_gpu"] = bool(flag.item() > 0.5)
    out["checked_at_nodes"] = nc
    del sh, full, want
    # ---- timing ----
    sh = ShardedAPSP(ctx, n_nodes, rank=rank, world=world, group=group)
    r0, r1 = sh.t0 * TILE, min(n_nodes, sh.t1 * TILE)
    d_rows = synthetic_cost_rows(n_nodes, r0, r1, dev)
    for tag, la in (("ms_lookahead", True), ("ms_no_lookahead", False)):
        if world == 1 and not la:
            continue
        best = None
        for _ in range(reps):
            sh.load_rows(d_rows)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier(group=group)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sh.run(lookahead=la)
            e1.record()
            torch.cuda.synchronize()
            ms = max_over_ranks(e0.elapsed_time(e1))
            best = ms if best is None else min(best, ms)
        out[tag] = best
    out["ms"] = out["ms_lookahead"]
    out["trelax_per_s"] = float(n_nodes) ** 3 / (out["ms"] * 1e-3) / 1e12
    out["rounds"] = sh.n_tiles
    out["panel_bytes"] = TILE * sh.n_pad * 4
    return out
