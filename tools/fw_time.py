"""K5 timings (CUDA events): dense random symmetric costs (bench configs[2] form) and a sparse
contact-like graph, FP32 / FP64, with and without predecessors.  python tools/fw_time.py 2000 10000"""
import json, sys, torch, numpy as np
sys.path.insert(0, '.')
from wisp_mdanalysis_implementation_b200 import _lib, synth
ctx = _lib.default_context(0)
out = {}
for n in [int(a) for a in sys.argv[1:]] or [2000, 10000]:
    g = torch.Generator(device='cuda'); g.manual_seed(1003)
    W = torch.rand((n, n), dtype=torch.float64, device='cuda', generator=g) * 2 + 0.05
    W = torch.triu(W, 1); W = W + W.T
    d = torch.empty((n, n), dtype=torch.float64, device='cuda')
    p = torch.empty((n, n), dtype=torch.int32, device='cuda')
    res = {}
    for prec in (32, 64):
        if prec == 64 and n > 12000:
            continue
        for with_pred in (False, True):
            pp = p.data_ptr() if with_pred else None
            ctx.dev_apsp(W.data_ptr(), n, prec, d.data_ptr(), pp, stream=1)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3 if n <= 4000 else 1
            e0.record()
            for _ in range(reps):
                ctx.dev_apsp(W.data_ptr(), n, prec, d.data_ptr(), pp, stream=1)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            res["fp%d%s" % (prec, "_pred" if with_pred else "")] = {"ms": round(ms, 3), "trelax_per_s": round(n ** 3 / ms / 1e9, 3)}
    out[str(n)] = res
print(json.dumps(out))
