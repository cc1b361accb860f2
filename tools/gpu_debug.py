"""Bisect helper: run one named step on the GPU with flush-prints (use under `timeout`)."""
import sys
import time
import numpy as np
import torch
from wisp_mdanalysis_implementation_b200 import _lib, synth

step = sys.argv[1]
n, t, apn = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (50, 100, 3)
ctx = _lib.default_context()
system, coords = synth.synth(n, t, apn, seed=7 + n)
print("step", step, n, t, apn, flush=True)
t0 = time.time()
if step == "com":
    nodes, avg = ctx.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
elif step in ("cov", "contact", "gram3", "dev"):
    nodes, avg = ctx.com_nodes(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx)
    print("com ok", flush=True)
    if step == "cov":
        ctx.cartesian_covariance(nodes, avg)
    elif step == "contact":
        ctx.contact_map(nodes, 15.0)
    else:
        ld = _lib.node_ld(n); tp = _lib.frames_pad(t); npad = _lib.gram_ld(n)
        x = torch.zeros(tp, ld, dtype=torch.float64, device="cuda")
        x[:t, :3 * n] = torch.from_numpy((nodes - avg).reshape(t, 3 * n)).cuda()
        g = torch.zeros(npad, npad, dtype=torch.float64, device="cuda")
        torch.cuda.synchronize()
        print("launch gram3", flush=True)
        ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g.data_ptr())
        ctx.sync()
        print("gram3 done", flush=True)
        ref = np.einsum("tic,tjc->ij", nodes - avg, nodes - avg)
        got = g.cpu().numpy()[:n, :n]
        iu = np.triu_indices(n)
        print("gram3 max rel err", np.abs(got[iu] - ref[iu]).max() / np.abs(ref).max(), flush=True)
        if step == "dev":
            mean = torch.zeros(ld, dtype=torch.float64, device="cuda")
            mean[:3 * n] = torch.from_numpy(avg.reshape(-1)).cuda()
            d = torch.zeros(npad, npad, dtype=torch.float64, device="cuda")
            torch.cuda.synchronize()
            p32 = torch.zeros(t, npad // 128, 3, 128, dtype=torch.float32, device="cuda")
            ctx.dev_contact_sum(p32.data_ptr(), t, n, mean.data_ptr(), d.data_ptr())
            ctx.sync()
            print("contact done", flush=True)
elif step == "pipeline":
    out = ctx.pipeline(coords, system.masses, system.node_ptr, system.atom_idx, system.align_idx,
                       cutoff=15.0, which_map=1, want_nodes=True)
print("done %.2fs" % (time.time() - t0), flush=True)
