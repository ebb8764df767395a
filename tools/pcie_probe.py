"""Pinned host<->device copy bandwidth of the box (context for the e2e number: 752x480 frames are 361 KB each)."""
import torch, time
n = 185 * 1024 * 1024
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
for name, src, dst in (("H2D", h, d), ("D2H", d, h)):
    for _ in range(3): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10): dst.copy_(src, non_blocking=True)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
    print("%s %.1f GB/s (%.2f ms per 185 MiB)" % (name, n / dt / 1e9, dt * 1e3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print("bidirectional: %.1f GB/s each way" % (n / dt / 1e9))
