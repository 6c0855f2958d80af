import torch, time
n = 224 * 1024 * 1024
d = torch.empty(n, dtype=torch.uint8, device="cuda")
h = torch.empty(n, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def one():
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
def two():
    k = n // 2
    with torch.cuda.stream(s1):
        h[:k].copy_(d[:k], non_blocking=True)
    with torch.cuda.stream(s2):
        h[k:].copy_(d[k:], non_blocking=True)
for name, fn in (("one stream", one), ("two streams", two), ("one stream", one), ("two streams", two)):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(5): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 5
    print(f"{name}: {dt*1e3:.2f} ms per 224 MiB = {n/dt/1e9:.1f} GB/s")
