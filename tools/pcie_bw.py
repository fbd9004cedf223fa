import torch, time
n = 504_000_000 // 4
h = torch.empty(n, dtype=torch.float32).pin_memory(); d = torch.empty(n, dtype=torch.float32, device="cuda")
for name, fn in (("H2D", lambda: d.copy_(h, non_blocking=True)), ("D2H", lambda: h.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(name, "%.1f GB/s  %.1f ms per 504 MB" % (n * 4 / dt / 1e9, dt * 1e3))
