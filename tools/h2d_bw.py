import torch, time
n = 4 << 30
h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
d = torch.empty(n, dtype=torch.uint8, device="cuda")
for _ in range(2):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D pinned GB/s: %.1f" % (3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
e0.record()
for _ in range(3):
    h.copy_(d, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("D2H pinned GB/s: %.1f" % (3 * n / (e0.elapsed_time(e1) * 1e-3) / 1e9))
