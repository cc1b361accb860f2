"""Time K1 (dev_com with a plan) alone on random coordinates: python tools/k1_time.py N T atoms_per_node reps"""
import sys
import torch
from wisp_mdanalysis_implementation_b200 import _lib, synth
from wisp_mdanalysis_implementation_b200.pipeline import FramePipeline

n, t, apn = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 5
ctx = _lib.default_context()
system = synth.SynthSystem(n, 8, apn, seed=1)
A = n * apn
pipe = FramePipeline(ctx, n, A, system.masses, system.node_ptr, system.atom_idx, system.align_idx, t)
gen = torch.Generator(device="cuda").manual_seed(2)
d_coords = (torch.rand(t, A, 3, generator=gen, device="cuda", dtype=torch.float32) - 0.5) * 80.0
fn = lambda: pipe.com(d_coords)
fn(); torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("K1 N=%d T=%d apn=%d: %.3f ms  (%.0f GB/s algorithmic)  checksum %.12e" %
      (n, t, apn, ms, t * (12.0 * A + 24 * n) / ms / 1e6, float(pipe.d_x[:t].sum())), flush=True)
