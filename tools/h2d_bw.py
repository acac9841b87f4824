"""pinned H2D bandwidth: one stream vs the same bytes split over several copy streams"""
import torch, time
mb = 16
h = torch.empty(mb * 2**20 // 4, dtype=torch.float32).pin_memory()
d = torch.empty_like(h, device="cuda")
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    hs, ds = h.chunk(ns), d.chunk(ns)
    def go():
        for s, a, b in zip(streams, hs, ds):
            with torch.cuda.stream(s):
                b.copy_(a, non_blocking=True)
    for _ in range(3): go()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    n = 50
    for _ in range(n): go()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print("H2D %d MB over %d streams: %.1f GB/s (%.0f us)" % (mb, ns, mb * 2**20 / dt / 1e9, dt * 1e6))
