"""Raw pinned-memory PCIe bandwidth of this box (sets the floor of bench.py's e2e number)."""
import torch, time
n = 64 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it
a = t(lambda: d.copy_(h, non_blocking=True)); b = t(lambda: h.copy_(d, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both)
print(f"H2D 64 MiB {a*1e3:.3f} ms = {n/a/1e9:.1f} GB/s; D2H {b*1e3:.3f} ms = {n/b/1e9:.1f} GB/s; both at once {c*1e3:.3f} ms = {2*n/c/1e9:.1f} GB/s aggregate")
