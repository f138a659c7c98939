import torch, time
n = 732 * 1024 * 1024 // 4
d = torch.empty(n, device="cuda"); h = torch.empty(n).pin_memory()
for name, fn in (("D2H", lambda: h.copy_(d, non_blocking=True)), ("H2D", lambda: d.copy_(h, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.2f} ms for 732 MiB = {n * 4 / ms / 1e6:.1f} GB/s")
