import torch, time
x = torch.empty(1<<30, dtype=torch.uint8, pin_memory=True)
d = torch.empty(1<<30, dtype=torch.uint8, device="cuda")
for name, (dst, src) in {"h2d": (d, x), "d2h": (x, d)}.items():
    for _ in range(2):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(4):
        dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    print(name, 4 * (1<<30) / (time.perf_counter() - t) / 1e9, "GB/s")
