import torch, time
n = 86*1024*1024
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, k=10):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(k): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / k
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"pcie 86MB: h2d {n/a/1e9:.1f} GB/s ({a*1e3:.2f} ms)  d2h {n/b/1e9:.1f} GB/s ({b*1e3:.2f} ms)  duplex {c*1e3:.2f} ms")
