"""Is host<->device copy bandwidth full duplex on this box? (development aid)"""
import torch, time
n = 1 << 30
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_a = torch.empty(n, dtype=torch.uint8, device="cuda")
d_b = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(3):
        if h2d:
            with torch.cuda.stream(s1): d_a.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_b, non_blocking=True)
    torch.cuda.synchronize(); return (time.perf_counter() - t) / 3
for name, a, b in (("h2d", 1, 0), ("d2h", 0, 1), ("both", 1, 1)):
    run(a, b); dt = run(a, b)
    print("%s: %.1f ms  (%.1f GB/s per direction)" % (name, dt * 1e3, n / dt / 1e9))
