#!/usr/bin/env python
"""bench.py -- WISP hot path on B200: corr+contact throughput in N^2*T/s (BASELINE.json).

    python bench.py --gpus 1 --steps K --warmup W            # this repository's CUDA path
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (port)
    torchrun ... bench.py --gpus N ...                       # one rank per GPU, NCCL

Workload (configs[1] of BASELINE.json): 2,000 nodes x 16 atoms, 50,000 frames, residue-COM
nodes, binary contact map at 15 A; synthetic trajectory (synth.py model, generated on the
device).  A "step" = one pass coordinates -> cost matrix: K1 node COM (+ alignment landmarks),
K7 iterative alignment (the reference always runs it, trajectory_analysis_functions.py:96),
centring, K2 Gram (tcgen05 INT8 digit planes / FP64 DMMA), K3 pair distances, (all-reduce), K4
epilogue.  Frames are sharded over ranks (strong scaling: the 50,000 frames are split).
``value`` has the coordinates resident in HBM; ``e2e`` is the ONE call a drop-in user makes --
wisp_pipeline through the C ABI on a context over all N GPUs, page-locked host coordinates in,
result matrices in host buffers out.  After the timed loops the results are CHECKED: a 4,096-frame
prefix goes through the same pipelines and is compared with the CPU oracle (``parity``).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "corr+contact node-pair-frames per second (N^2*T/s)"
UNIT = "pair-frames/s"


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "source_hbm": "fallback (B200_PROFILING.md)"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            mp = json.load(open(p))
            peaks["hbm_gbs"] = float(mp["hbm_gbs"])
            peaks["bf16_tflops"] = float(mp.get("bf16_tflops", 1633.1))
            peaks["source_hbm"] = "MEASURED_PEAKS.json"
        except Exception:
            pass
    # FP64 / FP32 / MUFU pipe peaks are not in MEASURED_PEAKS.json; measured on this pool's
    # B200 with tools/peaks.cu (register-resident loops) and tools/dgemm_peak.py (cuBLAS)
    q = os.path.join(ROOT, "profiles", "peaks_r1.json")
    peaks.update({"fp64_dmma_tflops": 37.08, "fp32_fma_tflops": 72.23, "mufu_sqrt_tops": 4.637,
                  "cublas_dgemm_tflops": 35.47, "source_pipes": "builtin (profiles/peaks_r1.json)"})
    if os.path.exists(q):
        try:
            peaks.update(json.load(open(q)))
            peaks["source_pipes"] = "profiles/peaks_r1.json (tools/peaks.cu on this pool's B200)"
        except Exception:
            pass
    # dispatch ceiling of a 32-bit min-plus relaxation (tools/peaks_minplus.cu, round 2)
    peaks["minplus_dispatch_ceiling_trelax"] = 17.9
    try:
        peaks["minplus_dispatch_ceiling_trelax"] = float(json.load(open(os.path.join(
            ROOT, "profiles", "peaks_minplus_r2.json")))["minplus_dispatch_ceiling_trelax"])
    except Exception:
        pass
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None,
                "power_w_max": float(max(power)) if power else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ---------------------------------------------------------------------------------------
# CPU arms (the oracle is the thing being TIMED here, never part of the product path)
# ---------------------------------------------------------------------------------------
def cpu_sample(system, frames, cutoff, literal=True, fixed_stages=True):
    """Time the oracle (port of the reference's CPU path) on `frames` frames of the workload.
    Returns (per-frame-stage seconds, T-independent seconds, detail dict).  The T-independent
    N^2 stages (Pearson trace loop, -log) can be skipped with fixed_stages=False."""
    import oracle
    coords = system.frames(0, frames)
    t0 = time.perf_counter()
    nodes, align = oracle.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    t_com = time.perf_counter() - t0
    t0 = time.perf_counter()
    nodes, avg = oracle.iterative_alignment(align, nodes)           # K7's stage: always runs in the reference
    t_align = time.perf_counter() - t0
    t0 = time.perf_counter()
    cov = oracle.cartesian_covariance(nodes, avg) if literal else oracle.cartesian_covariance_fast(nodes, avg)
    t_cov = time.perf_counter() - t0
    t0 = time.perf_counter()
    if literal:
        binary, avgdist = oracle.calc_contact_map(nodes, cutoff)
    else:
        binary, avgdist = oracle.calc_contact_map_fast(nodes, cutoff)
    t_contact = time.perf_counter() - t0
    t_pearson = t_func = 0.0
    if fixed_stages:
        t0 = time.perf_counter()
        if literal:
            _, _, corr = oracle.pearson_correlation_analysis(cov)
        else:
            _, _, corr = oracle.pearson_correlation_fast(cov)
        t_pearson = time.perf_counter() - t0
        t0 = time.perf_counter()
        oracle.func_pearson_correlation(corr, binary)
        t_func = time.perf_counter() - t0
    detail = dict(com_s=t_com, alignment_s=t_align, covariance_s=t_cov, pearson_s=t_pearson, contact_s=t_contact,
                  functionalize_s=t_func)
    return t_com + t_align + t_cov + t_contact, t_pearson + t_func, detail


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"]
        return max(n) if n else 1
    except Exception:
        return os.cpu_count() or 1


REFERENCE_BLAS_THREADS = 16      # the same at every N (torchrun exports OMP_NUM_THREADS=1 otherwise)


def run_reference_arm(args):
    """--impl reference: the reference's CPU path (oracle port, literal loop structure, numpy BLAS
    pinned to REFERENCE_BLAS_THREADS threads at every N) on bounded samples of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=min(REFERENCE_BLAS_THREADS, os.cpu_count() or 1))
    except Exception:
        pass
    from wisp_mdanalysis_implementation_b200 import synth
    system = synth.SynthSystem(args.nodes, max(args.sample_frames, 8), args.atoms_per_node, seed=1002)
    N, Ts, T = args.nodes, args.sample_frames, args.frames
    times = []
    detail = {}
    fixed = None
    for i in range(args.warmup + args.steps):
        # the N^2-only stages (Pearson trace loop, -log) run once per job, not once per frame:
        # time them once and amortise them over the full T like the real job would
        per_frame, fx, d = cpu_sample(system, Ts, args.cutoff, literal=True, fixed_stages=fixed is None)
        if fixed is None:
            fixed, detail = fx, d
        else:
            detail.update({k: v for k, v in d.items() if k in ("com_s", "alignment_s", "covariance_s", "contact_s")})
        t = per_frame + fixed * (Ts / T)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * float(np.mean(times))
    value = N * N * Ts / (ms * 1e-3)
    cores = blas_threads()
    sample = ("%d of %d frames, literal loop structure (per-frame COM calls, iterative alignment, per-frame rank-1 "
              "np.dot, per-pair np.trace, per-frame N x N distances); N^2-only stages amortised over T; BLAS "
              "threads pinned to %d at every N" % (Ts, T, cores))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": sample, "detail_s": detail},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "configs[1]: %d nodes x %d atoms, %d frames, residue-COM nodes, binary contact "
                        "map at %.1f A" % (args.nodes, args.atoms_per_node, args.frames, args.cutoff),
            "nodes": args.nodes, "frames": args.frames, "atoms": args.nodes * args.atoms_per_node,
            "paths": args.paths, "sharding": "frames over ranks",
            "alignment": "iterative alignment to the running average, threshold 1e-5, <= 100 iterations "
                         "(the reference's defaults); the synthetic trajectory has no global tumbling",
            "l2": "inputs larger than L2 (coordinate block per rank >= 2.4 GB vs 126 MB)"}


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=2000)
    ap.add_argument("--frames", type=int, default=50000)
    ap.add_argument("--atoms-per-node", type=int, default=16)
    ap.add_argument("--cutoff", type=float, default=15.0)
    ap.add_argument("--paths", type=int, default=500)
    ap.add_argument("--sample-frames", type=int, default=16, help="frames per CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-paths", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-paths-cap", type=int, default=20, help="seconds allowed for the CPU get_paths comparison")
    ap.add_argument("--align-iterations", type=int, default=100, help="K7 iteration limit (reference default 100; 0 = off)")
    ap.add_argument("--parity-frames", type=int, default=4096, help="frames of the prefix the results are checked on (0 = skip)")
    ap.add_argument("--parity-nodes", type=int, default=256, help="nodes of the oracle comparison")
    ap.add_argument("--stress-pairs", type=int, default=200)
    ap.add_argument("--stress-paths", type=int, default=1000)
    ap.add_argument("--apsp-nodes", type=int, default=10000, help="configs[2] APSP size (0 = skip)")
    ap.add_argument("--apsp-sharded-nodes", type=int, default=30000, help="configs[3] sharded APSP size (0 = skip)")
    args = ap.parse_args()

    if args.impl == "reference":
        run_reference_arm(args)
        return

    # stdout must carry exactly ONE JSON line: libraries (NCCL's version banner, ...) that
    # write to fd 1 are pointed at stderr for the whole run; the line goes to the saved fd.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    import torch
    import torch.distributed as dist
    from wisp_mdanalysis_implementation_b200 import _lib, synth
    from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline, bind_to_gpu_numa_node

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa_node = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        # a barrier that leaves the GPUs alone: while rank 0 drives all N GPUs through the in-process
        # wisp_pipeline call, the other ranks must wait on the CPU (an NCCL barrier is a spinning kernel
        # on every other GPU and gets time-sliced against the call's own kernels)
        cpu_group = dist.new_group(backend="gloo")
    dev = torch.device("cuda", local_rank)
    ctx = _lib.default_context(local_rank)

    N, T, apn = args.nodes, args.frames, args.atoms_per_node
    A = N * apn
    t_begin = T * rank // world
    t_end = T * (rank + 1) // world
    T_local = t_end - t_begin
    system = synth.SynthSystem(N, T, apn, seed=1002)

    # ---- synthetic coordinates generated on the device (no 19 GB host generation) ----
    d_coords = torch.empty((T_local, A, 3), dtype=torch.float32, device=dev)
    d_base = torch.as_tensor(system.atom_base, device=dev)
    d_noa = torch.as_tensor(system.node_of_atom, device=dev)
    d_modes = torch.as_tensor(np.ascontiguousarray(system.modes), device=dev)
    d_coeff = torch.as_tensor(np.ascontiguousarray(system.coeff), device=dev)
    cur = torch.cuda.current_stream(dev).cuda_stream or 1   # 0x1 == cudaStreamLegacy
    ctx.dev_synth_coords(d_coords.data_ptr(), t_begin, T_local, A, d_base.data_ptr(), d_noa.data_ptr(),
                         d_modes.data_ptr(), synth.N_MODES, d_coeff.data_ptr(), system.noise_sigma,
                         system.seed, stream=cur)
    torch.cuda.synchronize()

    pipe = FramePipeline(ctx, N, A, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                         frames_local=T_local, frames_total=T, cutoff=args.cutoff, which_map=1,
                         align_iterations=args.align_iterations)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident steps: `value` ----
    launches0 = ctx.kernel_launches()
    for _ in range(args.warmup):
        pipe.run_device(d_coords)
    barrier()
    launches_per_step = (ctx.kernel_launches() - launches0) // max(1, args.warmup)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    l0 = ctx.kernel_launches()
    e0.record()
    for _ in range(args.steps):
        pipe.run_device(d_coords)
    e1.record()
    barrier()
    gpu_launches = ctx.kernel_launches() - l0
    ms_step = max_over_ranks(e0.elapsed_time(e1) / args.steps)
    value = float(N) * N * T / (ms_step * 1e-3)

    # ---- per-kernel timing inside the same process (CUDA events on the launching stream) ----
    def time_stage(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    reps = max(2, args.steps)
    # order matters: K1 rewrites X (un-centred); mean_and_center centres it in place again
    aligned = args.align_iterations > 0
    k1_ms = time_stage(lambda: pipe.com(d_coords, colsum=not aligned), reps)
    k7 = None
    if aligned:
        torch.cuda.synchronize()
        a7, b7 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a7.record()
        n_it = pipe.align()
        b7.record()
        torch.cuda.synchronize()
        k7_ms = a7.elapsed_time(b7)
        # per frame and iteration: landmarks read (solve), landmarks read + written (float32), nodes READ
        # (float64; their rotation is composed and applied once, inside the centring pass), the new column
        # sums come out of the same pass; the initial (float32-chained) sums read the landmarks once (the
        # node sums came out of K1)
        n_al = len(system.align_idx)
        k7_bytes = T_local * ((n_it * 3 + 1) * (12.0 * n_al) + n_it * (24.0 * N))
        k7 = {"ms": k7_ms, "iterations": n_it, "ms_per_iteration": k7_ms / max(1, n_it), "bound": "hbm",
              "achieved_gbs": k7_bytes / (k7_ms * 1e-3) / 1e9, "peak_gbs": load_peaks()["hbm_gbs"],
              "frac": k7_bytes / (k7_ms * 1e-3) / 1e9 / load_peaks()["hbm_gbs"],
              "residuals": [r[1] for r in pipe.alignment_log],
              "bytes_per_frame_iteration": 36.0 * n_al + 24.0 * N,
              "note": "nodes are only read per iteration (rotations composed, applied once by the centring pass); "
                      "host loop: one all-reduce of 3(n_align+N) doubles and one 16-byte D2H per iteration"}
        pipe.center_on_mean()
    else:
        pipe.mean_and_center(have_colsum=True)
    k2_ms = time_stage(pipe.gram, reps)                       # K2 as the step runs it (auto engine)
    g_auto = pipe.d_sums[0].clone()
    ctx.set_gram_mode("dmma")
    k2_dmma_ms = time_stage(pipe.gram, reps)                  # FP64 DMMA engine, for comparison
    k2_engine = "fp64 dmma" if torch.equal(g_auto, pipe.d_sums[0]) else "int8 tcgen05"
    k2_i8_vs_dmma = float(((g_auto - pipe.d_sums[0]).abs().triu() /
                           torch.sqrt(torch.outer(pipe.d_sums[0].diagonal(), pipe.d_sums[0].diagonal())).clamp_min(1e-300)).max())
    ctx.set_gram_mode("auto")
    pipe.gram()
    del g_auto
    k3_ms = time_stage(pipe.contact, reps)
    k4_ms = time_stage(pipe.reduce_and_epilogue, reps)
    clocks = sampler.stop() if rank == 0 else None
    work = pipe.work()
    peaks = load_peaks()
    k2_tflops = work["k2_flop_executed"] / (k2_ms * 1e-3) / 1e12
    k2_dmma_tflops = work["k2_flop_executed"] / (k2_dmma_ms * 1e-3) / 1e12
    # INT8 engine: 10 digit-pair MMAs per FP64 product
    k2_int8_tops = 10.0 * k2_tflops if k2_engine == "int8 tcgen05" else 0.0
    # INT8 tcgen05 peak: the MMA kernel alone sustains 3,044 TOP/s with ncu reading the utcimma pipe at 72.9 %
    # (profiles/ncu_final_r1.json) -> 4,175 TOP/s at 100 %; 2 x the measured dense bf16 (3,266) would flatter
    int8_peak = peaks.get("int8_tcgen05_tops", 4175.0)
    k3_rate = work["k3_pairframes_executed"] / (k3_ms * 1e-3)
    k1_gbs = work["k1_bytes"] / (k1_ms * 1e-3) / 1e9
    kernels = {
        "k1_com": {"ms": k1_ms, "bound": "hbm", "achieved_gbs": k1_gbs, "peak_gbs": peaks["hbm_gbs"],
                   "frac": k1_gbs / peaks["hbm_gbs"]},
        "k2_gram": {"ms": k2_ms, "engine": k2_engine, "fp64_equivalent_tflops": k2_tflops,
                    "int8_tops_executed": k2_int8_tops, "int8_peak_tops": int8_peak,
                    "frac_int8": k2_int8_tops / int8_peak,
                    "max_abs_dev_vs_dmma_rel_sqrt_GiiGjj": k2_i8_vs_dmma,
                    "note": "ms covers the whole K2 stage: node statistics, digit planes, tcgen05 Gram, "
                            "partial reduce, plus the gated (empty) DMMA launch"},
        "k2_gram_dmma": {"ms": k2_dmma_ms, "bound": "tensor(fp64)", "achieved_tflops": k2_dmma_tflops,
                         "peak_tflops": peaks["fp64_dmma_tflops"],
                         "frac": k2_dmma_tflops / peaks["fp64_dmma_tflops"],
                         "frac_of_cublas_dgemm": k2_dmma_tflops / peaks["cublas_dgemm_tflops"],
                         # the same pass with K2 pinned to the FP64 engine (WISP_K2_MODE=dmma)
                         "pass_ms_with_this_engine": ms_step - k2_ms + k2_dmma_ms,
                         "value_with_this_engine": float(N) * N * T / ((ms_step - k2_ms + k2_dmma_ms) * 1e-3)},
        "k3_contact": {"ms": k3_ms, "bound": "mufu/fp32", "achieved_tpairframes": k3_rate / 1e12,
                       "mufu_peak_tops": peaks["mufu_sqrt_tops"],
                       "frac_mufu": k3_rate / 1e12 / peaks["mufu_sqrt_tops"],
                       "fp32_flop_frac": 8 * k3_rate / 1e12 / peaks["fp32_fma_tflops"]},
        "k4_epilogue_allreduce": {"ms": k4_ms},
    }
    if k7:
        kernels["k7_align"] = k7
    dominant = "k2_gram_dmma" if (k2_ms >= k3_ms and k2_engine == "fp64 dmma") else "k3_contact"
    if dominant == "k2_gram_dmma":
        roofline = {"kernel": "gram_dmma_kernel<3>", "bound": "tensor", "achieved": k2_tflops,
                    "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                    "frac": k2_tflops / peaks["fp64_dmma_tflops"], "traffic": None,
                    "peak_source": "FP64 DMMA pipe peak measured with tools/peaks.cu on this pool "
                                   "(MEASURED_PEAKS.json has no FP64 figure); cuBLAS DGEMM 8192^3 = %.2f"
                                   % peaks["cublas_dgemm_tflops"]}
    else:
        roofline = {"kernel": "contact_sum_ws_kernel", "bound": "fp32/mufu pipe", "achieved": k3_rate / 1e12,
                    "peak": peaks["mufu_sqrt_tops"], "unit": "Tsqrt/s", "frac": k3_rate / 1e12 / peaks["mufu_sqrt_tops"],
                    "traffic": None,
                    "peak_source": "MUFU.SQRT issue peak measured with tools/peaks.cu (1 sqrt per pair-frame); the same "
                                   "instruction mix held in registers tops out at %.2f T pair-frames/s" % peaks.get("k3_mix_tpairframes", 4.42)}

    if k2_engine == "int8 tcgen05":
        roofline_tensor = {"kernel": "gram_i8_kernel", "bound": "tensor", "achieved": k2_int8_tops, "peak": int8_peak,
                           "unit": "TOP/s", "frac": k2_int8_tops / int8_peak, "traffic": None,
                           "peak_source": "ncu-derived: gram_i8_kernel alone = 3,044 TOP/s at 72.9 % utcimma pipe "
                                          "utilisation -> 4,175 TOP/s (2 x measured dense bf16 would be 3,266); "
                                          "achieved counts the whole K2 stage time, not the MMA kernel alone"}
    else:
        roofline_tensor = {"kernel": "gram_dmma_kernel<3>", "bound": "tensor", "achieved": k2_tflops,
                           "peak": peaks["fp64_dmma_tflops"], "unit": "TFLOP/s",
                           "frac": k2_tflops / peaks["fp64_dmma_tflops"], "traffic": None}
    # the HBM-bound stage of the path (K1 = landmark gather + streaming COM), in the contract's own terms
    roofline_hbm = {"kernel": "align_shift_kernel + com_stream_kernel", "bound": "hbm", "achieved": k1_gbs,
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": k1_gbs / peaks["hbm_gbs"], "traffic": None,
                    "peak_source": peaks["source_hbm"]}
    try:
        # DRAM bytes per launch from the newest committed `ncu --set full` capture (tools/ncu_summarize.py)
        import glob
        traffic_file = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_traffic_r*.json")))[-1]
        ncu = json.load(open(traffic_file))
        roofline["traffic"] = ncu.get(roofline["kernel"].split("<")[0])
        roofline_tensor["traffic"] = ncu.get(roofline_tensor["kernel"].split("<")[0])
        roofline_hbm["traffic"] = ncu.get("com_stream_kernel", 0.0) + ncu.get("align_shift_kernel", 0.0)
        for rf in (roofline, roofline_tensor, roofline_hbm):
            rf["traffic_source"] = "profiles/%s: %s" % (os.path.basename(traffic_file), ncu.get(
                "_source", "ncu --set full of the round-1 build; these kernels are unchanged since"))
    except Exception:
        pass

    # ---- e2e: the ONE call a drop-in user makes.  wisp_pipeline (C ABI) on a context over all N GPUs of the
    # box, page-locked host coordinates in, result matrices in page-locked host buffers out.  Under torchrun the
    # other ranks wait at a barrier while rank 0 makes the call: the library shards the frames over the N GPUs
    # itself (one host thread, stream and PCIe link per GPU; partial sums combined over NVLink peer memory inside
    # the epilogue kernel).  The per-rank NCCL path (FramePipeline.run_host) is timed beside it. ----
    e2e = None
    if not args.no_e2e:
        def pinned(shape, dtype):
            return torch.empty(shape, dtype=dtype, pin_memory=True)

        # (a) per-rank path: every rank feeds its own frames from its own pinned buffer, NCCL all-reduce
        h_local = pinned((T_local, A, 3), torch.float32)
        h_local.copy_(d_coords)
        torch.cuda.synchronize()
        # barrier-aligned concurrent H2D probe: every rank copies 2 GiB x 3 at the same time
        nprobe = min(T_local, max(1, (2 << 30) // (A * 12)))
        probe = torch.empty((nprobe, A, 3), dtype=torch.float32, device=dev)
        probe.copy_(h_local[:nprobe], non_blocking=True)
        barrier()
        pa, pb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pa.record()
        for _ in range(3):
            probe.copy_(h_local[:nprobe], non_blocking=True)
        pb.record()
        torch.cuda.synchronize()
        probe_ms = pa.elapsed_time(pb)
        h2d_gbs_rank = 3 * nprobe * A * 12 / (probe_ms * 1e-3) / 1e9
        h2d_gbs_aggregate = world * 3 * nprobe * A * 12 / (max_over_ranks(probe_ms) * 1e-3) / 1e9
        del probe
        for _ in range(2):
            pipe.run_host(h_local)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(args.steps):
            pipe.run_host(h_local)
        b.record()
        barrier()
        ms_rank_path = max_over_ranks(a.elapsed_time(b) / args.steps)

        # (b) the C-ABI call on rank 0 over all N GPUs
        ms_call, wall_call, api_gpus = None, None, world
        host_nodes = None
        if world > 1:
            # gather the whole trajectory into ONE page-locked buffer on rank 0 (setup, untimed); frame
            # shard r is first-touched on the NUMA node of GPU r (pipeline.host_trajectory_numa_sharded)
            if rank == 0:
                affinity0 = os.sched_getaffinity(0)
                os.sched_setaffinity(0, range(os.cpu_count()))       # undo the per-rank binding for the threads of the call

                def fill(r, view):
                    dst = torch.from_numpy(view)
                    if r == 0:
                        dst.copy_(d_coords)
                    else:
                        tmp = torch.empty(view.shape, dtype=torch.float32, device=dev)
                        dist.recv(tmp, src=r)
                        dst.copy_(tmp)
                        del tmp
                from wisp_mdanalysis_implementation_b200.pipeline import host_trajectory_numa_sharded
                h_np, host_nodes = host_trajectory_numa_sharded((T, A, 3), world, fill)
            else:
                dist.send(d_coords, dst=0)
            torch.cuda.synchronize()
        else:
            h_np = h_local.numpy()
        cpu_barrier()
        if rank == 0:
            mctx = _lib.Context(gpus=world) if world > 1 else ctx
            mctx_ingest = mctx.ingest_info()
            outbuf = {"avg": pinned((N, 3), torch.float64).numpy(), "corr": pinned((N, N), torch.float64).numpy(),
                      "avgdist": pinned((N, N), torch.float64).numpy(), "binary": pinned((N, N), torch.int64).numpy(),
                      "w": pinned((N, N), torch.float64).numpy()}

            def call():
                return mctx.pipeline(h_np, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                                     cutoff=args.cutoff, which_map=1, maximum_num_iterations=args.align_iterations,
                                     out=outbuf)
            for _ in range(max(3, args.warmup)):      # W >= 3 untimed calls (scratch growth, lazy module loads, host pages)
                res_call = call()
            torch.cuda.synchronize()
            ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0 = time.perf_counter()
            ea.record()
            call_ms, call_stages = [], []
            for _ in range(args.steps):
                tc = time.perf_counter()
                res_call = call()          # blocking: returns when every GPU's slice is in the host buffers
                call_ms.append((time.perf_counter() - tc) * 1e3)
                call_stages.append({k: round(v, 2) for k, v in mctx.pipeline_stats().items()})
            eb.record()
            torch.cuda.synchronize()
            wall_call = (time.perf_counter() - t0) * 1e3 / args.steps
            ms_call = ea.elapsed_time(eb) / args.steps
            e2e_corr = res_call["corr"].copy()
            e2e_avgdist = res_call["avgdist"].copy()
            e2e_iters = len(res_call["alignment_log"])
            direct_ms = None
            if any(v != i for i, v in enumerate(mctx_ingest["via"])):
                # the same call with every GPU pulling its own shard over its own host link, for comparison
                os.environ["WISP_INGEST"] = "direct"
                dctx = _lib.Context(gpus=world)
                os.environ.pop("WISP_INGEST")
                keep, mctx = mctx, dctx
                call()
                td = time.perf_counter()
                for _ in range(3):
                    call()
                direct_ms = (time.perf_counter() - td) * 1e3 / 3
                mctx = keep
                dctx.close()
        cpu_barrier()
        if rank == 0:
            h2d_total = T * A * 12
            e2e = {"value": float(N) * N * T / (ms_call * 1e-3), "unit": UNIT, "ms_per_step": ms_call,
                   "wall_ms_per_step": wall_call, "wall_ms_each_call": call_ms,
                   "wall_ms_median_call": float(np.median(call_ms)),
                   "stages_of_each_call_ms": call_stages, "h2d_bytes_per_step": h2d_total,
                   "d2h_bytes_per_step": 4 * N * N * 8 + N * 24,
                   "api": "wisp_pipeline (C ABI, include/wisp_b200.h) on wisp_ctx_create_multi(%d): page-locked host "
                          "coordinates in; avg / corr / avgdist / binary / w in page-locked host buffers out; "
                          "K7 alignment included (%d iterations)" % (api_gpus, e2e_iters),
                   "gpus_in_call": api_gpus, "ingest": mctx_ingest, "ms_per_step_with_direct_ingest": direct_ms,
                   "host_numa_node_per_shard": host_nodes,
                   "pcie_h2d_gbs_per_rank_concurrent": h2d_gbs_rank,
                   "pcie_h2d_gbs_aggregate_concurrent": h2d_gbs_aggregate,
                   "frac_of_pcie_bound": (h2d_total / (h2d_gbs_aggregate * 1e9)) / (ms_call * 1e-3),
                   "per_rank_nccl_path": {"ms_per_step": ms_rank_path, "value": float(N) * N * T / (ms_rank_path * 1e-3),
                                          "api": "FramePipeline.run_host on every rank (wisp_dev_* C-ABI calls, "
                                                 "K7 all-reduces, one NCCL all-reduce of [G || D])"}}
            if world > 1:
                mctx.close()
                _lib.host_unregister(h_np)
                os.sched_setaffinity(0, affinity0)
        del h_local

    # ---- parity: the SAME pipelines on a frame prefix, checked against the CPU oracle on a node subset (every
    # matrix entry depends on its two nodes only), the multi-rank result against a single-GPU recomputation,
    # and the e2e call's full-length result against the device-resident pass ----
    parity = None
    if args.parity_frames > 0:
        Tp = int(min(args.parity_frames, T // world))
        d_prefix = d_coords[:Tp].clone() if rank == 0 else torch.empty((Tp, A, 3), dtype=torch.float32, device=dev)
        if world > 1:
            dist.broadcast(d_prefix, src=0)
        q0, q1 = Tp * rank // world, Tp * (rank + 1) // world
        ctx.set_gram_mode("i8" if k2_engine == "int8 tcgen05" else "dmma")      # the engine the timed pass used
        try:
            pp = FramePipeline(ctx, N, A, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                               frames_local=q1 - q0, frames_total=Tp, cutoff=args.cutoff, which_map=1,
                               align_iterations=args.align_iterations)
            pp.run_device(d_prefix[q0:q1].contiguous())
            torch.cuda.synchronize()
            if rank == 0:
                import oracle
                got = dict(corr=pp.d_corr.cpu().numpy(), avgdist=pp.d_avgdist.cpu().numpy(),
                           binary=pp.d_binary.cpu().numpy())
                parity = {"frames": Tp, "k2_engine_forced": "i8 (accuracy-gated)" if k2_engine == "int8 tcgen05" else "dmma",
                          "alignment_iterations": len(pp.alignment_log)}
                if world > 1:
                    one = FramePipeline(ctx, N, A, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                                        frames_local=Tp, frames_total=Tp, cutoff=args.cutoff, which_map=1,
                                        align_iterations=args.align_iterations, group=False)
                    one.run_device(d_prefix)
                    torch.cuda.synchronize()
                    c1, d1, b1 = one.d_corr.cpu().numpy(), one.d_avgdist.cpu().numpy(), one.d_binary.cpu().numpy()
                    parity["multi_rank_vs_single_gpu"] = {
                        "ranks": world,
                        "max_rel_corr": float((np.abs(got["corr"] - c1) / np.maximum(np.abs(c1), 1e-3)).max()),
                        "max_rel_dist": float((np.abs(got["avgdist"] - d1) / np.maximum(d1, 1e-9))[~np.eye(N, dtype=bool)].max()),
                        "topology_mismatches": int((got["binary"] != b1).sum())}
                    del one
                # oracle on a node subset that touches every tile class
                rng = np.random.default_rng(7)
                sel = np.unique(np.concatenate([rng.choice(N, size=min(N, args.parity_nodes), replace=False),
                                                [0, min(N - 1, 127), min(N - 1, 128), N - 1]]))
                ptr = np.concatenate([[0], np.cumsum(np.diff(system.node_ptr)[sel])]).astype(np.int64)
                idx = np.concatenate([system.atom_idx[system.node_ptr[i]:system.node_ptr[i + 1]] for i in sel])
                h_prefix = d_prefix.cpu().numpy()
                t0 = time.perf_counter()
                nodes_o, align_o = oracle.com_nodes_fast(h_prefix, system.masses, ptr, idx, system.align_idx)
                if args.align_iterations > 0:
                    nodes_o, avg_o = oracle.iterative_alignment(align_o, nodes_o, maximum_num_iterations=args.align_iterations)
                else:
                    avg_o = nodes_o.sum(axis=0) / Tp
                _, _, corr_o = oracle.pearson_correlation_fast(oracle.cartesian_covariance_fast(nodes_o, avg_o))
                bin_o, dist_o = oracle.calc_contact_map_fast(nodes_o, args.cutoff)
                ix = np.ix_(sel, sel)
                off = ~np.eye(len(sel), dtype=bool)
                rel_c = np.abs(got["corr"][ix] - corr_o) / np.maximum(np.abs(corr_o), 1e-3)
                rel_d = (np.abs(got["avgdist"][ix] - dist_o) / np.maximum(dist_o, 1e-9))[off]
                near = np.abs(dist_o - args.cutoff) <= 1e-6 * args.cutoff
                mism = int(((got["binary"][ix] != bin_o) & ~near).sum())
                parity.update({"nodes_checked": int(len(sel)), "entries_checked": int(len(sel)) ** 2,
                               "max_rel_corr": float(rel_c.max()), "max_rel_dist": float(rel_d.max()),
                               "topology_mismatches": mism, "oracle_s": time.perf_counter() - t0,
                               "tolerance": "1e-6 relative (|C| floor 1e-3), topology exact outside |D - cutoff| <= 1e-6 cutoff",
                               "pass": bool(rel_c.max() <= 1e-6 and rel_d.max() <= 1e-6 and mism == 0)})
                if e2e is not None:
                    # the e2e call worked on ALL frames: compare with the device-resident pass (same data, same stages)
                    parity["e2e_call_vs_resident_pass"] = {
                        "max_abs_corr": float(np.abs(e2e_corr - pipe.d_corr.cpu().numpy()).max()),
                        "max_rel_dist": float((np.abs(e2e_avgdist - pipe.d_avgdist.cpu().numpy())
                                               / np.maximum(e2e_avgdist, 1e-9))[~np.eye(N, dtype=bool)].max())}
            del pp
        finally:
            ctx.set_gram_mode("auto")
        del d_prefix
        barrier()

    # ---- APSP + 500 suboptimal paths on the resulting cost matrix (rank 0) ----
    extra = {}
    if rank == 0 and not args.no_paths:
        W = pipe.d_w.cpu().numpy()
        hop_w = (W != 0).astype(np.float64)
        t0 = time.perf_counter()
        hops, _ = ctx.apsp(hop_w, precision=32, want_pred=False)
        t_hops = time.perf_counter() - t0
        cand = np.argwhere((hops >= 5) & (hops <= 8))
        cand = cand[cand[:, 0] < cand[:, 1]]
        s, t = (int(cand[0][0]), int(cand[0][1])) if len(cand) else (0, N - 1)
        def time_apsp(d_W, n, prec, with_pred):
            d_dist = torch.empty((n, n), dtype=torch.float64, device=dev)
            d_pred = torch.empty((n, n), dtype=torch.int32, device=dev) if with_pred else None
            pp = d_pred.data_ptr() if with_pred else None
            ctx.dev_apsp(d_W.data_ptr(), n, prec, d_dist.data_ptr(), pp, stream=cur)
            torch.cuda.synchronize()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
            ctx.dev_apsp(d_W.data_ptr(), n, prec, d_dist.data_ptr(), pp, stream=cur)
            ev1.record()
            torch.cuda.synchronize()
            return ev0.elapsed_time(ev1)

        d_W = torch.as_tensor(W, device=dev)
        peaks_ = load_peaks()
        for prec, with_pred, tag in ((64, True, "fp64"), (32, True, "fp32"), (32, False, "fp32_dist_only")):
            ms = time_apsp(d_W, N, prec, with_pred)
            extra["apsp_%s_ms" % tag] = ms
            extra["apsp_%s_trelax_per_s" % tag] = float(N) ** 3 / (ms * 1e-3) / 1e12
        # configs[2]: 10,000-node single-GPU blocked Floyd-Warshall (dense random symmetric costs)
        if args.apsp_nodes > 0:
            n3 = args.apsp_nodes
            g = torch.Generator(device=dev)
            g.manual_seed(1003)
            Wb = torch.rand((n3, n3), dtype=torch.float64, device=dev, generator=g) * 2.0 + 0.05
            Wb = torch.triu(Wb, 1)
            Wb = Wb + Wb.T
            big = {}
            for prec, with_pred, tag in ((32, False, "fp32_dist_only"), (32, True, "fp32"), (64, True, "fp64")):
                ms = time_apsp(Wb, n3, prec, with_pred)
                rate = float(n3) ** 3 / (ms * 1e-3) / 1e12
                big[tag] = {"ms": ms, "trelax_per_s": rate,
                            "frac_of_fp32_pipe_peak": 2.0 * rate / peaks_["fp32_fma_tflops"],
                            # pipe bound of a relaxation: adds on the FP32 pipe (FADD2) vs mins on the ALU pipe
                            "frac_of_add_min_pipe_bound": rate / min(peaks_.get("fp32_ffma2_tflops", 72.1) / 2.0,
                                                                     peaks_.get("fmnmx_tops", 36.6)),
                            # two dispatch cycles per relaxation whatever the instruction form (FADD2+FMNMX3,
                            # FADD+FMNMX, DPX VIADDMNMX): 17.9 T/s = 49.6 % of the FP32 peak is the machine's limit
                            "frac_of_dispatch_ceiling": rate / peaks_["minplus_dispatch_ceiling_trelax"]}
            extra["apsp_%d_nodes" % n3] = big
            del Wb
        t0 = time.perf_counter()
        paths = ctx.get_paths(W, [s], [t], args.paths)
        extra["get_paths_wall_ms"] = (time.perf_counter() - t0) * 1e3
        extra["get_paths"] = dict(source=s, sink=t, returned=len(paths), shortest=paths[0][0],
                                  longest=paths[-1][0], **ctx.paths_stats())
        extra["hops_apsp_wall_ms"] = t_hops * 1e3
        # ---- the north-star run in one number: page-locked host coordinates -> wisp_pipeline on all N GPUs
        # (K1, K7, K2, K3, fused cross-GPU K4) -> APSP (FP64, predecessors) + 500 suboptimal paths, with the
        # matrices checked against the oracle on the frame prefix (parity) and the paths checked here ----
        if e2e is not None:
            import scipy.sparse
            import scipy.sparse.csgraph
            ok = len(paths) == args.paths or len(paths) == args.paths + 1
            lens = [p[0] for p in paths]
            ok &= all(lens[i] <= lens[i + 1] for i in range(len(lens) - 1))
            seen = set()
            for p in paths:
                nodes = p[1:]
                lr = 0.0
                for a, b in zip(nodes[:-1], nodes[1:]):
                    ok &= W[a, b] != 0
                    lr += W[a, b]
                ok &= nodes[0] == s and nodes[-1] == t and len(set(nodes)) == len(nodes) and lr == p[0]
                ok &= tuple(nodes) not in seen
                seen.add(tuple(nodes))
            d_ref = scipy.sparse.csgraph.dijkstra(scipy.sparse.csr_matrix(W), directed=True, indices=[s])[0]
            ok &= abs(paths[0][0] - d_ref[t]) <= 1e-12 * max(1.0, d_ref[t])
            extra["north_star"] = {
                "what": "coordinates (page-locked host memory) -> correlation + contacts on %d GPU(s) -> APSP -> %d "
                        "suboptimal paths" % (world, args.paths),
                "pipeline_call_ms": e2e["ms_per_step"], "apsp_and_paths_ms": extra["get_paths_wall_ms"],
                "total_ms": e2e["ms_per_step"] + extra["get_paths_wall_ms"],
                "matrices_checked": "parity (oracle on the %d-frame prefix)" % args.parity_frames,
                "matrices_pass": bool(parity["pass"]) if parity else None,
                "paths_checked": "count K or K+1, sorted, simple, every hop an edge, LR float64 lengths bit-exact, "
                                 "no duplicates, shortest == scipy Dijkstra",
                "paths_pass": bool(ok)}
        # configs[4]: 200 (source, sink) pairs x 1,000 suboptimal paths each, one APSP
        if args.stress_pairs > 0:
            rng = np.random.default_rng(1005)
            cand5 = np.argwhere((hops >= 4) & (hops <= 10))
            cand5 = cand5[cand5[:, 0] < cand5[:, 1]]
            sel = cand5[rng.choice(len(cand5), size=min(args.stress_pairs, len(cand5)), replace=False)]
            t0 = time.perf_counter()
            ctx.paths_prepare(W)
            t_prep = time.perf_counter() - t0
            # all pairs in ONE batched call: the cutoff passes of every unfinished query share the kernels
            t0 = time.perf_counter()
            qoff, plen, poff, pnodes = ctx.paths_query_pairs(sel, args.stress_paths, packed=True)
            t_q = time.perf_counter() - t0
            st_b = ctx.paths_stats()
            # one pair at a time (round 1's form), on a subset, for the comparison
            sub = sel[:20]
            t0 = time.perf_counter()
            n_sub = 0
            for (a, b) in sub:
                n_sub += len(ctx.paths_query([int(a)], [int(b)], args.stress_paths))
            t_seq = time.perf_counter() - t0
            same = True
            for k, (a, b) in enumerate(sub[:5]):
                one = ctx.paths_query([int(a)], [int(b)], args.stress_paths)
                lo, hi = int(qoff[k]), int(qoff[k + 1])
                same &= (hi - lo == len(one)) and all(
                    plen[lo + j] == one[j][0] and pnodes[poff[lo + j]:poff[lo + j + 1]].tolist() == one[j][1:]
                    for j in range(hi - lo))
            extra["stress_c5"] = {"pairs": int(len(sel)), "paths_each": args.stress_paths,
                                  "prepare_ms": t_prep * 1e3, "queries_wall_ms": t_q * 1e3,
                                  "api": "wisp_paths_query_pairs (one batched call)",
                                  "paths_returned": int(len(plen)), "expansions": int(st_b["expansions"]),
                                  "expansions_per_s": st_b["expansions"] / max(t_q, 1e-9),
                                  "kernel_rounds": int(st_b["final_cutoff"]), "enumerate_ms": st_b["enumerate_ms"],
                                  "host_ms": st_b["host_ms"],
                                  "one_pair_at_a_time_ms_per_pair": t_seq * 1e3 / max(1, len(sub)),
                                  "batched_equals_one_at_a_time_on_5_pairs": bool(same)}

    # ---- configs[4] over all ranks: pairs dealt round-robin, one APSP per rank, gather on rank 0 ----
    if world > 1 and not args.no_paths and args.stress_pairs > 0:
        from wisp_mdanalysis_implementation_b200 import paths_sharded
        W = pipe.d_w.cpu().numpy()              # identical on every rank after the all-reduce + K4
        hops, _ = ctx.apsp((W != 0).astype(np.float64), precision=32, want_pred=False)
        rng = np.random.default_rng(1005)
        cand5 = np.argwhere((hops >= 4) & (hops <= 10))
        cand5 = cand5[cand5[:, 0] < cand5[:, 1]]
        sel = cand5[rng.choice(len(cand5), size=min(args.stress_pairs, len(cand5)), replace=False)]
        barrier()
        t0 = time.perf_counter()
        res = paths_sharded.query_pairs(ctx, W, sel, args.stress_paths, packed=True)
        barrier()
        t_all = max_over_ranks((time.perf_counter() - t0) * 1e3)
        if rank == 0:
            t0 = time.perf_counter()
            first = res[0]                       # Python lists are built on demand, per query
            t_unpack = (time.perf_counter() - t0) * 1e3
            extra["stress_c5_sharded"] = {"ranks": world, "pairs": int(len(sel)), "paths_each": args.stress_paths,
                                          "wall_ms_incl_apsp_and_gather": t_all,
                                          "paths_returned": res.total_paths(),
                                          "unpack_one_query_to_lists_ms": t_unpack, "first_query_paths": len(first),
                                          "gather": "four packed arrays per rank (no per-path Python objects)"}

    # ---- configs[3]: 30,000-node row-block sharded Floyd-Warshall over all ranks (look-ahead), checked on
    # this hardware against the single-GPU matrix at 8,192 nodes ----
    if args.apsp_sharded_nodes > 0 and not args.no_paths:
        from wisp_mdanalysis_implementation_b200.apsp_sharded import run_sharded_benchmark
        del d_coords
        torch.cuda.empty_cache()
        resc4 = run_sharded_benchmark(ctx, args.apsp_sharded_nodes, rank, world)
        if rank == 0:
            resc4["frac_of_dispatch_ceiling"] = resc4["trelax_per_s"] / world / load_peaks()["minplus_dispatch_ceiling_trelax"]
            extra["apsp_sharded_c4"] = resc4

    # ---- CPU baseline (rank 0, N = 1 only): the oracle timed on the box's host cores ----
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        Ts = args.sample_frames
        per_frame, fixed, detail = cpu_sample(system, Ts, args.cutoff, literal=True)
        t_lit = per_frame + fixed * (Ts / T)
        per_frame_v, fixed_v, detail_v = cpu_sample(system, 4 * Ts, args.cutoff, literal=False)
        t_vec = per_frame_v + fixed_v * (4 * Ts / T)
        cpu_baseline = {"value": float(N) * N * Ts / t_lit, "unit": UNIT, "cores": blas_threads(),
                        "kind": "port",
                        "sample": "%d of %d frames through the oracle with the reference's loop structure; "
                                  "N^2-only stages amortised over T" % (Ts, T),
                        "detail_s": detail,
                        "vectorised": {"value": float(N) * N * 4 * Ts / t_vec,
                                       "sample": "%d frames, one GEMM + blocked distances (numpy, all BLAS threads)" % (4 * Ts),
                                       "detail_s": detail_v}}
        if "get_paths" in extra:
            # CPU shortest-path baselines on the same cost matrix: (L) the three pure-Python
            # Dijkstra calls the reference makes per (pair, pass) (network_analysis_functions.py:
            # 30,71-72), (V) scipy's C Dijkstra from every source = a CPU APSP
            import oracle
            import scipy.sparse
            import scipy.sparse.csgraph
            t0 = time.perf_counter()
            adj = oracle.graph_edges(W)
            oracle.dijkstra_path(W, s, t, adj)
            oracle.dijkstra(W, s, adj)
            oracle.dijkstra(W, t, adj)
            extra["cpu_three_dijkstra_python_s"] = time.perf_counter() - t0
            t0 = time.perf_counter()
            scipy.sparse.csgraph.dijkstra(scipy.sparse.csr_matrix(W), directed=False, return_predecessors=True)
            extra["cpu_apsp_scipy_dijkstra_s"] = time.perf_counter() - t0
            # the reference's get_paths on the same cost matrix through the pruned-equivalent
            # oracle, in a child process under a wall-clock cap (it is exponential; the literal
            # frontier loop would not finish at all -- BASELINE.md section 3)
            import tempfile
            cap_s = args.cpu_paths_cap
            with tempfile.TemporaryDirectory() as td:
                np.save(os.path.join(td, "w.npy"), W)
                code = ("import sys,time,numpy as np; sys.path.insert(0,%r); import oracle; "
                        "W=np.load(%r); t0=time.perf_counter(); "
                        "p=oracle.get_paths_pruned(W,[%d],[%d],%d); print(time.perf_counter()-t0)"
                        % (ROOT, os.path.join(td, "w.npy"), s, t, min(args.paths, 100)))
                try:
                    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True,
                                       timeout=cap_s)
                    extra["cpu_get_paths_pruned_100_s"] = float(r.stdout.strip().splitlines()[-1])
                except subprocess.TimeoutExpired:
                    extra["cpu_get_paths_pruned_100_s"] = ">%d (capped; 288 s measured in round 1)" % cap_s
                except Exception as exc:   # pragma: no cover
                    extra["cpu_get_paths_pruned_100_s"] = "failed: %s" % exc

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "dtype_note": "results float64; K2 products as exact INT8 digit planes on tcgen05 with FP64 combine "
                              "(accuracy-gated, DMMA FP64 fallback), K3 per-frame distances FP32 with FP64 sums",
                "config": workload_config(args), "clocks": clocks, "e2e": e2e,
                "gpu_launches": int(gpu_launches), "launches_per_step": int(launches_per_step),
                "roofline": roofline, "roofline_tensor": roofline_tensor, "roofline_hbm": roofline_hbm, "kernels": kernels, "parity": parity, "cpu_baseline": cpu_baseline, "extra": extra}
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
