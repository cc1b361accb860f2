"""Time K3 (dev_contact_sum) alone on random tile-referenced data: python tools/k3_time.py N T reps"""
import sys
import torch
from wisp_mdanalysis_implementation_b200 import _lib

n, t = int(sys.argv[1]), int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
ctx = _lib.default_context()
ld = _lib.node_ld(n); npad = _lib.gram_ld(n)
gen = torch.Generator(device="cuda").manual_seed(1)
p32 = (torch.rand(t, npad // 128, 3, 128, generator=gen, device="cuda", dtype=torch.float32) - 0.5) * 40.0
mean = torch.zeros(ld, dtype=torch.float64, device="cuda")
mean[:3 * n] = (torch.rand(3 * n, generator=gen, device="cuda", dtype=torch.float64) - 0.5) * 100.0
out = torch.zeros(npad, npad, dtype=torch.float64, device="cuda")
fn = lambda: ctx.dev_contact_sum(p32.data_ptr(), t, n, mean.data_ptr(), out.data_ptr(), stream=1)
fn(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
tiles = npad // 128
print("K3 N=%d T=%d: %.3f ms  (%.3f T pair-frames/s executed)  checksum %.6e" %
      (n, t, ms, tiles * (tiles + 1) / 2 * 16384 * t / ms / 1e9, float(out.sum())), flush=True)
