"""Host->device bandwidth of this box, one process, one thread per GPU (the topology wisp_pipeline on a
multi-GPU context uses): every GPU alone, all at once, subsets that reveal shared PCIe switches / sockets,
default page-locked buffers vs buffers first-touched on the GPU's NUMA node.  Prints one JSON object.

    python tools/h2d_probe.py [n_gpus]
"""
import json
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisp_mdanalysis_implementation_b200.pipeline import gpu_numa_node  # noqa: E402

GIB = 1 << 30
BYTES = 2 * GIB
REPS = 3


def alloc(dev, local):
    """2 GiB page-locked buffer; local: first-touched (cudaHostAlloc touches) on the GPU's NUMA node."""
    before = os.sched_getaffinity(0)
    node, cpus = gpu_numa_node(dev)
    try:
        if local and cpus:
            os.sched_setaffinity(0, cpus)
        h = torch.empty(BYTES // 4, dtype=torch.float32, pin_memory=True)
        h.fill_(1.0)
    finally:
        os.sched_setaffinity(0, before)
    return h


def run(devs, bufs):
    """Concurrent copies on `devs`; returns (per-GPU GB/s, aggregate GB/s)."""
    bar = threading.Barrier(len(devs))
    out, t_start, t_end = {}, {}, {}

    def work(d):
        torch.cuda.set_device(d)
        dst = torch.empty(BYTES // 4, dtype=torch.float32, device="cuda:%d" % d)
        st = torch.cuda.Stream(device=d)
        with torch.cuda.stream(st):
            dst.copy_(bufs[d], non_blocking=True)
            st.synchronize()
            bar.wait()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t_start[d] = time.perf_counter()
            e0.record(st)
            for _ in range(REPS):
                dst.copy_(bufs[d], non_blocking=True)
            e1.record(st)
            st.synchronize()
            t_end[d] = time.perf_counter()
        out[d] = REPS * BYTES / (e0.elapsed_time(e1) * 1e-3) / 1e9

    th = [threading.Thread(target=work, args=(d,)) for d in devs]
    for t in th:
        t.start()
    for t in th:
        t.join()
    agg = len(devs) * REPS * BYTES / (max(t_end.values()) - min(t_start.values())) / 1e9
    return {str(d): round(out[d], 1) for d in devs}, round(agg, 1)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
    n = min(n, torch.cuda.device_count())
    res = {"gpus": n, "cpus": os.cpu_count(), "numa": {}}
    for d in range(n):
        node, cpus = gpu_numa_node(d)
        res["numa"][str(d)] = {"node": node, "n_cpus": len(cpus) if cpus else None}
    try:
        res["nodes_online"] = open("/sys/devices/system/node/online").read().strip()
    except Exception:
        res["nodes_online"] = None
    try:
        res["topo"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
    except Exception as exc:
        res["topo"] = "unavailable: %s" % exc
    default = {d: alloc(d, False) for d in range(n)}
    res["alone_default"] = {str(d): run([d], default)[0][str(d)] for d in range(n)}
    per, agg = run(list(range(n)), default)
    res["all_default"] = {"per_gpu": per, "aggregate": agg}
    subsets = [s for s in ([0, 1], [0, 2], [0, 4], [0, 1, 2, 3], [4, 5, 6, 7], [0, 2, 4, 6], [0, 4, 1, 5]) if max(s) < n]
    res["subsets_default"] = {}
    for sset in subsets:
        per, agg = run(sset, default)
        res["subsets_default"][",".join(map(str, sset))] = {"per_gpu": per, "aggregate": agg}
    del default
    local = {d: alloc(d, True) for d in range(n)}
    per, agg = run(list(range(n)), local)
    res["all_numa_local"] = {"per_gpu": per, "aggregate": agg}
    if n >= 4:
        per, agg = run(list(range(4)), local)
        res["first4_numa_local"] = {"per_gpu": per, "aggregate": agg}
    # one shared buffer read by every GPU (what a caller-owned trajectory on one node looks like)
    shared = {d: local[0] for d in range(n)}
    per, agg = run(list(range(n)), shared)
    res["all_one_buffer_on_node_of_gpu0"] = {"per_gpu": per, "aggregate": agg}
    print(json.dumps(res))


if __name__ == "__main__":
    main()
