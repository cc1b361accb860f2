"""Frame-sharded device pipeline: coordinates -> cost matrix on 1..8 B200s.

One process per GPU.  Frames are the data-parallel axis: every rank runs
K1 (translate + node COM) on its own frame block, the ranks agree on the global mean with a
tiny all-reduce of the column sums, each rank centres its block and runs K2 (DMMA Gram) and
K3 (pair distances) on it, and the partial ``[G || D]`` sums -- ONE contiguous buffer -- are
combined by a single NCCL all-reduce before the fused epilogue K4.  torch is used only for
device memory, streams and ``torch.distributed``; every kernel is this repository's own,
called through the C ABI with raw pointers.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, sharding


def gpu_numa_node(device_index):
    """(NUMA node, set of its CPUs) of the PCIe root the GPU hangs off, or (None, None)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:          # 00000000:1b:00.0 -> 0000:1b:00.0
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        return node, cpus
    except Exception:
        return None, None


def bind_to_gpu_numa_node(device_index):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off, so that pinned host
    buffers allocated afterwards are node-local and every rank's H2D stream crosses only its own
    PCIe root (matters at 8 ranks: 8 x 50 GB/s of host reads).  Best effort; returns the node or
    None."""
    import os
    node, cpus = gpu_numa_node(device_index)
    if node is None:
        return None
    try:
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        return None
    return None


def host_trajectory_numa_sharded(shape, n_shards, fill):
    """Page-locked float32 host array of `shape` = (frames, atoms, 3) whose frame shard r (the frames
    wisp_pipeline on an n_shards-GPU context gives GPU r) is physically placed on the NUMA node of
    GPU r: the pages are first touched -- by ``fill(r, view)`` writing the shard -- while the calling
    thread is bound to that node's CPUs, then the whole range is page-locked in place
    (wisp_host_register).  Returns (array, [node per shard]); call _lib.host_unregister when done."""
    import os
    arr = np.empty(shape, dtype=np.float32)
    T = shape[0]
    before = os.sched_getaffinity(0)
    nodes = []
    try:
        for r in range(n_shards):
            node, cpus = gpu_numa_node(r)
            nodes.append(node)
            if cpus:
                try:
                    os.sched_setaffinity(0, cpus)
                except Exception:
                    pass
            fill(r, arr[T * r // n_shards:T * (r + 1) // n_shards])
    finally:
        os.sched_setaffinity(0, before)
    _lib.host_register(arr)
    return arr, nodes


class FramePipeline:
    def __init__(self, ctx, n_nodes, n_atoms, masses, node_ptr, atom_idx, align_idx,
                 frames_local, frames_total=None, cutoff=15.0, which_map=1, atomic=False,
                 group=None, align_iterations=0, align_threshold=1e-5):
        self.ctx = ctx
        self.dev = torch.device("cuda", ctx.device)
        self.N, self.A = int(n_nodes), int(n_atoms)
        self.T = int(frames_local)
        self.T_total = int(frames_total if frames_total is not None else frames_local)
        self.cutoff, self.which_map, self.atomic = float(cutoff), int(which_map), bool(atomic)
        self.group = group
        # K7 (trajectory_analysis_functions.py:86-126): 0 = translation only (pre-aligned input)
        self.align_iterations, self.align_threshold = int(align_iterations), float(align_threshold)
        self.n_align = len(align_idx)
        self.alignment_log = []
        self.ld = _lib.node_ld(self.N)
        self.tpad = _lib.frames_pad(self.T)
        self.npad = _lib.gram_ld(self.N)
        dev = self.dev
        f64, i32 = torch.float64, torch.int32
        self.plan = ctx.com_plan(self.A, masses, node_ptr, atom_idx, align_idx, atomic=self.atomic)
        self.plan_info = ctx.com_plan_info(self.plan)
        self.d_x = torch.zeros((self.tpad, self.ld), dtype=f64, device=dev)
        self.d_p32 = torch.zeros((self.T, self.npad // 128, 3, 128), dtype=torch.float32, device=dev)
        self.d_colsum = torch.zeros(self.ld, dtype=f64, device=dev)
        self.d_mean = torch.zeros(self.ld, dtype=f64, device=dev)
        # ONE contiguous buffer [G || D || s]: Gram sums, distance sums, and (streamed mode) the
        # column sums of the provisionally centred trajectory -> one all-reduce
        nn = self.npad * self.npad
        self.d_flat = torch.zeros(2 * nn + self.ld, dtype=f64, device=dev)
        self.d_sums = self.d_flat[:2 * nn].view(2, self.npad, self.npad)
        self.d_offs = self.d_flat[2 * nn:]
        self.d_ref = torch.zeros(self.ld, dtype=f64, device=dev)
        if self.align_iterations > 0:
            self.d_align = torch.empty((self.T, self.n_align, 3), dtype=torch.float32, device=dev)
            self.d_align_sums = torch.zeros(3 * self.n_align + 3 * self.N, dtype=f64, device=dev)
            # total rotation of every frame (K7 composes, the centring pass applies it once)
            self.d_rot_total = torch.empty((self.T, 9), dtype=f64, device=dev)
            self.d_node_sums = torch.zeros(self.ld, dtype=f64, device=dev)
        self._rotation_pending = False
        self._align_host = None
        self.d_parts = None
        self.k23_stream = torch.cuda.Stream(device=dev)
        n = self.N
        self.d_corr = torch.empty((n, n), dtype=f64, device=dev)
        self.d_avgdist = torch.empty((n, n), dtype=f64, device=dev)
        self.d_binary = torch.empty((n, n), dtype=torch.int64, device=dev)
        self.d_w = torch.empty((n, n), dtype=f64, device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.chunk_frames = max(1, min(self.T, (128 << 20) // (12 * self.A)))
        self._stage = None
        self.timings = {}

    # -- stages (all on torch's current stream) ------------------------------------------
    def _stream(self):
        # 0 is torch's legacy default stream; the C ABI reads NULL as "the context's own
        # stream", so name the legacy stream explicitly (cudaStreamLegacy == 0x1)
        return torch.cuda.current_stream(self.dev).cuda_stream or 1

    def com(self, d_coords, frames=None, row0=0, colsum=False):
        """K1 on a coordinate block; colsum=True also leaves the column sums of the rows just
        written in d_colsum (fused into the streaming kernel)."""
        t = self.T if frames is None else frames
        d_al = self.d_align.data_ptr() + row0 * self.n_align * 12 if self.align_iterations > 0 else None
        # with K7 on, the node column sums K1 produces seed the alignment's initial node average
        # (accumulated over the coordinate blocks of a streamed pass)
        want_sums = colsum or self.align_iterations > 0
        self.ctx.dev_com_align(self.plan, d_coords.data_ptr(), t, self.N,
                               self.d_x.data_ptr() + row0 * self.ld * 8,
                               self.d_colsum.data_ptr() if want_sums else None, d_al, stream=self._stream())
        if self.align_iterations > 0:
            if row0 == 0:
                self.d_node_sums.copy_(self.d_colsum)
            else:
                self.d_node_sums.add_(self.d_colsum)

    def align(self):
        """K7, frame-sharded: every rank rotates its frames onto the current landmark average; the
        new [landmark || node] column sums are all-reduced (3 (n_align + N) doubles per iteration)
        and every rank takes the same convergence decision.  Leaves the frame mean of the aligned
        node trajectory in d_mean (the node average of the last iteration IS that mean) and
        returns the number of iterations."""
        s = self._stream()
        na3, nn3 = 3 * self.n_align, 3 * self.N
        world = torch.distributed.get_world_size(self.group) if sharding.is_distributed(self.group) else 1
        # initial landmark totals (:86): sequential float32 sums, chained over the ranks in frame order --
        # rank q continues from the totals rank q-1 left -- so the result is the single sequential sum.
        # The node sums came out of K1 (com()).
        self.d_align_sums.zero_()
        for q in range(world):
            if q == (torch.distributed.get_rank(self.group) if world > 1 else 0):
                self.ctx.dev_align_sums(self.d_align.data_ptr(), self.n_align, None, self.T, self.N,
                                        True, self.d_align_sums.data_ptr(), stream=s)
            if world > 1:
                src = torch.distributed.get_global_rank(self.group, q) if self.group is not None else q
                torch.distributed.broadcast(self.d_align_sums[:na3], src=src, group=self.group)
        node_sums = self.d_node_sums[:nn3].clone()
        if world > 1:
            torch.distributed.all_reduce(node_sums, group=self.group)
        # averages stay on the device ([landmarks || nodes], the layout of the sums), formed by device-side
        # tensor ops: nothing comes back to the host here.  The float32 mean of :86 is a float32 division by
        # a DEVICE tensor (torch turns a division by a host scalar into a multiplication by its reciprocal,
        # which rounds differently).
        t32 = torch.full((1,), float(self.T_total), dtype=torch.float32, device=self.dev)
        t64 = torch.full((1,), float(self.T_total), dtype=torch.float64, device=self.dev)
        d_avg = torch.cat(((self.d_align_sums[:na3].to(torch.float32) / t32).to(torch.float64), node_sums / t64))
        # The loop decision lives on the device: align_update raises d_gate once the residual is below the
        # threshold, and every kernel of an iteration is a no-op while it is set.  The host therefore
        # launches iteration k + 1 BEFORE it has seen the residual of iteration k (one iteration of
        # look-ahead: the GPU never waits for the 24-byte D2H + the launch latency); when iteration k
        # turns out to be the last one, the already queued iteration k + 1 costs four empty launches.
        # Every rank derives the flag from the same all-reduced sums, so the ranks stay in step.
        d_res3 = torch.zeros(3, dtype=torch.float64, device=self.dev)
        d_gate = torch.zeros(1, dtype=torch.int32, device=self.dev)
        max_it = self.align_iterations
        if self._align_host is None or self._align_host.shape[0] < max_it:
            self._align_host = torch.empty((max_it, 3), dtype=torch.float64).pin_memory()
        h_res, events = self._align_host, []

        def launch(k):
            self.ctx.dev_align_iter(self.d_align.data_ptr(), self.n_align, d_avg.data_ptr(),
                                    self.d_x.data_ptr(), self.T, self.N, self.d_rot_total.data_ptr(), k == 0,
                                    d_sums=self.d_align_sums.data_ptr(), d_gate=d_gate.data_ptr(), stream=s)
            sharding.reduce_sums(self.d_align_sums, group=self.group)
            self.ctx.dev_align_update_gated(self.d_align_sums.data_ptr(), self.T_total, self.n_align, self.N,
                                            d_avg.data_ptr(), d_res3.data_ptr(), self.align_threshold,
                                            d_gate.data_ptr(), stream=s)
            h_res[k].copy_(d_res3, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.dev))
            events.append(ev)

        self.alignment_log = []
        it = 0
        if max_it > 0:
            launch(0)
        while it < max_it:
            if it + 1 < max_it:
                launch(it + 1)                      # speculative: gated off on the device if `it` converges
            events[it].synchronize()
            r2 = h_res[it]
            residual = float(np.sqrt(float(r2[0]) / self.n_align))
            node_rmsd = float(np.sqrt(float(r2[1]) / self.N))
            it += 1
            self.alignment_log.append((it, residual, node_rmsd))
            if not residual > self.align_threshold:
                break
        self._rotation_pending = it > 0
        self.d_mean.zero_()
        self.d_mean[:nn3].copy_(d_avg[na3:])
        return it

    def center_on_mean(self):
        """Centre with the mean already in d_mean (after align()); the node rotation the alignment
        deferred is applied in the same pass."""
        if self._rotation_pending:
            self.ctx.dev_center_rotated(self.d_x.data_ptr(), self.T, self.N, self.d_mean.data_ptr(),
                                        self.d_rot_total.data_ptr(), d_p32=self.d_p32.data_ptr(),
                                        stream=self._stream())
            self._rotation_pending = False
        else:
            self.ctx.dev_center(self.d_x.data_ptr(), self.T, self.N, self.d_mean.data_ptr(),
                                d_p32=self.d_p32.data_ptr(), stream=self._stream())

    def mean_and_center(self, have_colsum=False):
        s = self._stream()
        if not have_colsum:
            self.ctx.dev_colsum(self.d_x.data_ptr(), self.T, self.N, self.d_colsum.data_ptr(), stream=s)
        sharding.global_mean(self.d_colsum, self.d_mean, self.T_total, group=self.group)
        self.ctx.dev_center(self.d_x.data_ptr(), self.T, self.N, self.d_mean.data_ptr(),
                            d_p32=self.d_p32.data_ptr(), stream=s)

    def gram(self):
        self.ctx.dev_gram(self.d_x.data_ptr(), self.T, self.ld, self.N, 3,
                          self.d_sums[0].data_ptr(), stream=self._stream())

    def contact(self):
        self.ctx.dev_contact_sum(self.d_p32.data_ptr(), self.T, self.N, self.d_mean.data_ptr(),
                                 self.d_sums[1].data_ptr(), stream=self._stream())

    def reduce_and_epilogue(self):
        sharding.reduce_sums(self.d_sums, group=self.group)    # the ONE big collective
        self.ctx.dev_epilogue(self.d_sums[0].data_ptr(), self.d_sums[1].data_ptr(), self.T_total,
                              self.N, self.cutoff, self.which_map, d_corr=self.d_corr.data_ptr(),
                              d_avgdist=self.d_avgdist.data_ptr(),
                              d_binary=self.d_binary.data_ptr(), d_w=self.d_w.data_ptr(),
                              stream=self._stream())

    # -- whole steps ------------------------------------------------------------------
    def run_device(self, d_coords):
        """One pass with the coordinate block already resident in HBM."""
        if self.align_iterations > 0:
            self.com(d_coords)
            self.align()
            self.center_on_mean()
        else:
            self.com(d_coords, colsum=True)
            self.mean_and_center(have_colsum=True)
        self.gram_contact_reduce_epilogue()

    def gram_contact_reduce_epilogue(self):
        """K2, K3, the cross-rank sums and K4.  With several ranks the Gram all-reduce is issued on a side
        stream as soon as K2 is done and travels over NVLink while K3 (76 % of the pass) is computing;
        only the distance sums are reduced after K3."""
        if not sharding.is_distributed(self.group):
            self.gram()
            self.contact()
            self.reduce_and_epilogue()
            return
        main = torch.cuda.current_stream(self.dev)
        self.gram()
        done_k2 = torch.cuda.Event()
        done_k2.record(main)
        with torch.cuda.stream(self.k23_stream):
            self.k23_stream.wait_event(done_k2)
            sharding.reduce_sums(self.d_sums[0], group=self.group)
        self.contact()
        sharding.reduce_sums(self.d_sums[1], group=self.group)
        main.wait_stream(self.k23_stream)
        self.ctx.dev_epilogue(self.d_sums[0].data_ptr(), self.d_sums[1].data_ptr(), self.T_total,
                              self.N, self.cutoff, self.which_map, d_corr=self.d_corr.data_ptr(),
                              d_avgdist=self.d_avgdist.data_ptr(),
                              d_binary=self.d_binary.data_ptr(), d_w=self.d_w.data_ptr(),
                              stream=self._stream())

    def run_host(self, h_coords, fetch=True):
        """One pass from PINNED host coordinates (T_local, A, 3) float32: chunked H2D on a
        copy stream double-buffered against K1, then the device stages, then the results are
        read back to the host."""
        main = torch.cuda.current_stream(self.dev)
        if self._stage is None:
            self._stage = [torch.empty((self.chunk_frames, self.A, 3), dtype=torch.float32,
                                       device=self.dev) for _ in range(2)]
            self._h_out = [torch.empty((self.N, self.N), dtype=torch.float64).pin_memory()
                           for _ in range(3)] + [torch.empty((self.N, self.N), dtype=torch.int64).pin_memory()]
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        it = 0
        for t0 in range(0, self.T, self.chunk_frames):
            b = it & 1
            nt = min(self.chunk_frames, self.T - t0)
            with torch.cuda.stream(self.copy_stream):
                if it >= 2:
                    self.copy_stream.wait_event(freed[b])
                self._stage[b][:nt].copy_(h_coords[t0:t0 + nt], non_blocking=True)
                copied[b].record(self.copy_stream)
            main.wait_event(copied[b])
            self.com(self._stage[b], frames=nt, row0=t0)
            freed[b].record(main)
            it += 1
        if self.align_iterations > 0:
            self.align()
            self.center_on_mean()
        else:
            self.mean_and_center()
        self.gram_contact_reduce_epilogue()
        if fetch:
            self._h_out[0].copy_(self.d_corr, non_blocking=True)
            self._h_out[1].copy_(self.d_avgdist, non_blocking=True)
            self._h_out[2].copy_(self.d_w, non_blocking=True)
            self._h_out[3].copy_(self.d_binary, non_blocking=True)
            main.synchronize()
            return dict(corr=self._h_out[0].numpy(), avgdist=self._h_out[1].numpy(),
                        w=self._h_out[2].numpy(), binary=self._h_out[3].numpy())
        return None

    def run_host_streamed(self, h_coords, n_blocks=8, fetch=True):
        """Like run_host, but the frames are processed in blocks so that K2 / K3 of block b run
        (on a second stream) while block b+1 is still crossing PCIe.  No pass over the data
        waits for the global mean: every block is centred with a provisional reference r (the
        mean of the first block, broadcast from rank 0), and the exact parallel-axis term
        (s_i . s_j)/T with s = sum_t (x - r) is removed in the epilogue -- so the ranks exchange
        a SINGLE all-reduce of [G || D || s].

        Like run_host, returns numpy views of the pipeline's pinned output buffers: they are
        overwritten by the next run_host / run_host_streamed call -- copy what must outlive it.

        Only for pre-aligned input (align_iterations == 0): the alignment needs every frame before
        the first rotation is known, so nothing downstream of it can overlap the transfer."""
        if self.align_iterations > 0:
            raise ValueError("run_host_streamed cannot overlap K2/K3 with the transfer when K7 alignment is on; "
                             "use run_host")
        main = torch.cuda.current_stream(self.dev)
        side = self.k23_stream
        if self._stage is None:
            self._stage = [torch.empty((self.chunk_frames, self.A, 3), dtype=torch.float32,
                                       device=self.dev) for _ in range(2)]
            self._h_out = [torch.empty((self.N, self.N), dtype=torch.float64).pin_memory()
                           for _ in range(3)] + [torch.empty((self.N, self.N), dtype=torch.int64).pin_memory()]
        block = -(-self.T // n_blocks)
        block = max(24, -(-block // 24) * 24)
        starts = list(range(0, self.T, block))
        if self.d_parts is None or self.d_parts.shape[1] != len(starts):
            self.d_parts = torch.zeros((2, len(starts), self.npad, self.npad), dtype=torch.float64,
                                       device=self.dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]
        tile_floats = (self.npad // 128) * 384
        it = 0
        for b, f0 in enumerate(starts):
            f1 = min(self.T, f0 + block)
            for t0 in range(f0, f1, self.chunk_frames):
                k = it & 1
                nt = min(self.chunk_frames, f1 - t0)
                with torch.cuda.stream(self.copy_stream):
                    if it >= 2:
                        self.copy_stream.wait_event(freed[k])
                    self._stage[k][:nt].copy_(h_coords[t0:t0 + nt], non_blocking=True)
                    copied[k].record(self.copy_stream)
                main.wait_event(copied[k])
                self.com(self._stage[k], frames=nt, row0=t0)
                freed[k].record(main)
                it += 1
            if b == 0:
                s = self._stream()
                self.ctx.dev_colsum(self.d_x.data_ptr(), f1 - f0, self.N, self.d_colsum.data_ptr(), stream=s)
                torch.div(self.d_colsum, float(f1 - f0), out=self.d_ref)
                if sharding.is_distributed(self.group):
                    torch.distributed.broadcast(self.d_ref, src=0, group=self.group)
            done_k1 = torch.cuda.Event()
            done_k1.record(main)
            with torch.cuda.stream(side):
                side.wait_event(done_k1)
                ss = side.cuda_stream or 1
                nb = f1 - f0
                self.ctx.dev_center(self.d_x.data_ptr() + f0 * self.ld * 8, nb, self.N, self.d_ref.data_ptr(),
                                    d_p32=self.d_p32.data_ptr() + f0 * tile_floats * 4, stream=ss)
                self.ctx.dev_gram(self.d_x.data_ptr() + f0 * self.ld * 8, nb, self.ld, self.N, 3,
                                  self.d_parts[0, b].data_ptr(), stream=ss)
                self.ctx.dev_contact_sum(self.d_p32.data_ptr() + f0 * tile_floats * 4, nb, self.N,
                                         self.d_ref.data_ptr(), self.d_parts[1, b].data_ptr(), stream=ss)
        main.wait_stream(side)
        s = self._stream()
        self.ctx.dev_colsum(self.d_x.data_ptr(), self.T, self.N, self.d_offs.data_ptr(), stream=s)
        for k in (0, 1):
            self.ctx.dev_reduce_partials(self.d_parts[k].data_ptr(), len(starts), self.N,
                                         self.d_sums[k].data_ptr(), stream=s)
        sharding.reduce_sums(self.d_flat, group=self.group)          # the ONE collective
        self.ctx.dev_epilogue(self.d_sums[0].data_ptr(), self.d_sums[1].data_ptr(), self.T_total,
                              self.N, self.cutoff, self.which_map, d_corr=self.d_corr.data_ptr(),
                              d_avgdist=self.d_avgdist.data_ptr(), d_binary=self.d_binary.data_ptr(),
                              d_w=self.d_w.data_ptr(), d_offset_sum=self.d_offs.data_ptr(), stream=s)
        torch.add(self.d_ref, self.d_offs, alpha=1.0 / self.T_total, out=self.d_mean)
        if fetch:
            self._h_out[0].copy_(self.d_corr, non_blocking=True)
            self._h_out[1].copy_(self.d_avgdist, non_blocking=True)
            self._h_out[2].copy_(self.d_w, non_blocking=True)
            self._h_out[3].copy_(self.d_binary, non_blocking=True)
            main.synchronize()
            return dict(corr=self._h_out[0].numpy(), avgdist=self._h_out[1].numpy(),
                        w=self._h_out[2].numpy(), binary=self._h_out[3].numpy())
        return None

    def h2d_bytes(self):
        return self.T * self.A * 12

    def d2h_bytes(self):
        return 4 * self.N * self.N * 8

    # -- algorithmic work of one pass on this rank -----------------
    def work(self):
        n_tiles = self.npad // 128
        pairs = n_tiles * (n_tiles + 1) // 2
        return dict(
            k1_bytes=self.T * (12 * self.A + 24 * self.N),
            k2_flop_executed=2.0 * pairs * 128 * 128 * 3 * self.tpad,
            k2_flop_full=6.0 * self.N * self.N * self.T,
            k3_pairframes_full=float(self.N) * self.N * self.T,
            k3_pairframes_executed=pairs * 128.0 * 128.0 * self.T,
        )
