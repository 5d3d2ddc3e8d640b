import time, torch
dev = torch.device("cuda:0")
n = 1572864000
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
print("pinned:", h.is_pinned())
for name, fn in (("d2h", lambda: h.copy_(d, non_blocking=True)), ("h2d", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(); fn(); b.record(); torch.cuda.synchronize()
    print(name, "%.1f GB/s" % (n / a.elapsed_time(b) / 1e6))
# view-shaped like the bench
sol = torch.empty((1000, 65536, 3), dtype=torch.float64, device=dev)
oh = torch.empty((1000, 65536, 3), dtype=torch.float64).pin_memory()
oh.copy_(sol, non_blocking=True); torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); oh.copy_(sol, non_blocking=True); b.record(); torch.cuda.synchronize()
print("bench-shaped d2h %.1f GB/s" % (sol.numel() * 8 / a.elapsed_time(b) / 1e6))
