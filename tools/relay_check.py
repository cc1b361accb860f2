"""wisp_pipeline on an in-process multi-GPU context with the NVLink relay forced on ($WISP_INGEST=relay:
the second half of the GPUs is fed through the first half) against the direct path and a single GPU.
python tools/relay_check.py [n_gpus]   (needs >= 2 GPUs)"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wisp_mdanalysis_implementation_b200 import _lib, synth  # noqa: E402

n_gpus = int(sys.argv[1]) if len(sys.argv) > 1 else 2
system, coords = synth.synth(300, 6000, 16, seed=3)
_lib.host_register(coords)
one = _lib.Context(0).pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=13.0,
                               which_map=1, maximum_num_iterations=100)
out = {}
for mode in ("direct", "relay"):
    os.environ["WISP_INGEST"] = mode
    ctx = _lib.Context(gpus=n_gpus)
    info = ctx.ingest_info()
    t0 = time.perf_counter()
    got = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx, cutoff=13.0,
                       which_map=1, maximum_num_iterations=100)
    ms = (time.perf_counter() - t0) * 1e3
    e_c = float((np.abs(got["corr"] - one["corr"]) / np.maximum(np.abs(one["corr"]), 1e-3)).max())
    e_d = float(np.abs(got["avgdist"] - one["avgdist"]).max())
    same_topology = bool(np.array_equal(got["binary"], one["binary"]))
    out[mode] = {"via": info["via"], "max_rel_corr_vs_one_gpu": e_c, "max_abs_dist_vs_one_gpu": e_d,
                 "topology_equal": same_topology, "ms": ms}
    assert e_c <= 1e-6 and same_topology, (mode, e_c)
    if mode == "relay":
        assert any(v != i for i, v in enumerate(info["via"])), "relay was not engaged"
    ctx.close()
_lib.host_unregister(coords)
print(json.dumps(out))
