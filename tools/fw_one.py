"""One APSP call (for ncu): python tools/fw_one.py N precision with_pred"""
import sys, torch
sys.path.insert(0, '.')
from wisp_mdanalysis_implementation_b200 import _lib
n, prec, wp = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
ctx = _lib.default_context(0)
g = torch.Generator(device='cuda'); g.manual_seed(1003)
W = torch.rand((n, n), dtype=torch.float64, device='cuda', generator=g) * 2 + 0.05
W = torch.triu(W, 1); W = W + W.T
d = torch.empty((n, n), dtype=torch.float64, device='cuda')
p = torch.empty((n, n), dtype=torch.int32, device='cuda')
ctx.dev_apsp(W.data_ptr(), n, prec, d.data_ptr(), p.data_ptr() if wp else None, stream=1)
torch.cuda.synchronize()
print("ok")
