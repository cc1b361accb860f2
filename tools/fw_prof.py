import sys, torch, numpy as np
sys.path.insert(0, '.')
from wisp_mdanalysis_implementation_b200 import _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 6000
ctx = _lib.default_context(0)
g = torch.Generator(device='cuda'); g.manual_seed(1)
W = torch.rand((n, n), dtype=torch.float64, device='cuda', generator=g) * 2 + 0.05
W = torch.triu(W, 1); W = W + W.T
d = torch.empty((n, n), dtype=torch.float64, device='cuda')
for _ in range(2):
    ctx.dev_apsp(W.data_ptr(), n, 32, d.data_ptr(), None, stream=1)
torch.cuda.synchronize()
print("ok")
