"""INT8 tensor-core Gram vs the DMMA Gram on random data (run under `timeout`)."""
import sys
import numpy as np
import torch
from wisp_mdanalysis_implementation_b200 import _lib

n, t = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
ctx = _lib.default_context()
ld = _lib.node_ld(n); tp = _lib.frames_pad(t); npad = _lib.gram_ld(n)
gen = torch.Generator(device="cuda").manual_seed(5)
x = torch.zeros(tp, ld, dtype=torch.float64, device="cuda")
amp = torch.exp(torch.randn(3 * n, generator=gen, device="cuda", dtype=torch.float64))      # per-column spread
x[:t, :3 * n] = torch.randn(t, 3 * n, generator=gen, device="cuda", dtype=torch.float64) * amp
x[:t, :3 * n] -= x[:t, :3 * n].mean(dim=0, keepdim=True)
g0 = torch.zeros(npad, npad, dtype=torch.float64, device="cuda")
g1 = torch.full((npad, npad), -7.0, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
print("launch dmma", flush=True)
ctx.set_gram_mode("dmma")
ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g0.data_ptr(), stream=1)
torch.cuda.synchronize()
print("launch i8", flush=True)
ctx.dev_gram_i8(x.data_ptr(), t, ld, n, g1.data_ptr(), stream=1)
torch.cuda.synchronize()
print("i8 done", flush=True)
a = g0.cpu().numpy()[:n, :n]; b = g1.cpu().numpy()[:n, :n]
iu = np.triu_indices(n)
dg = np.sqrt(np.outer(np.diag(a), np.diag(a)))
err = np.abs(a[iu] - b[iu]) / dg[iu]
print("max |dG|/sqrt(GiiGjj) = %.3e   max rel = %.3e" % (err.max(), (np.abs(a[iu] - b[iu]) / np.maximum(np.abs(a[iu]), 1e-300)).max()), flush=True)
if err.max() > 1e-6:
    print("sample a", a[:3, :3], "b", b[:3, :3], sep="\n")
    bad = np.argwhere(np.abs(np.triu(a - b)) / dg > 1e-6)
    print("n bad", len(bad), "first", bad[:10].tolist())
if reps:
    ctx.set_gram_mode("auto")
    g2 = torch.zeros_like(g0)
    ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g2.data_ptr(), stream=1)
    torch.cuda.synchronize()
    print("auto == i8:", bool(torch.equal(torch.triu(g2[:n, :n]), torch.triu(g1[:n, :n]))), flush=True)
    x[7, 5] = 1e5 * float(x[:t, 5].abs().max())           # one spike: the accuracy gate must refuse
    ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g2.data_ptr(), stream=1)
    ctx.set_gram_mode("dmma")
    ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g0.data_ptr(), stream=1)
    torch.cuda.synchronize()
    print("spiky auto == dmma:", bool(torch.equal(torch.triu(g2[:n, :n]), torch.triu(g0[:n, :n]))), flush=True)
    def run_mode(mode):
        def f():
            ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g2.data_ptr(), stream=1)
        return mode, f
    for name, fn in (("dmma", lambda: ctx.dev_gram(x.data_ptr(), t, ld, n, 3, g0.data_ptr(), stream=1)),
                     ("i8", lambda: ctx.dev_gram_i8(x.data_ptr(), t, ld, n, g1.data_ptr(), stream=1))):
        fn(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        print("%s: %.3f ms" % (name, e0.elapsed_time(e1) / reps), flush=True)
