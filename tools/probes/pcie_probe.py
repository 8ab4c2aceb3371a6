"""Host <-> device copy bandwidth of this box (pinned memory): one direction at a time and both at once."""
import time
import torch
dev = torch.device("cuda", 0)
for mb in (46, 187, 1024):
    n = mb * (1 << 20)
    h = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    h2 = torch.empty(n, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=10):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    def d2h():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    def h2d():
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    def both():
        d2h(); h2d()
    def d2h_two():      # two D2H copies on two streams
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
        with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
    t = run(d2h); print("%5d MB  D2H %.1f GB/s" % (mb, n / t / 1e9))
    t = run(h2d); print("%5d MB  H2D %.1f GB/s" % (mb, n / t / 1e9))
    t = run(both); print("%5d MB  both %.1f GB/s each" % (mb, n / t / 1e9))
    t = run(d2h_two); print("%5d MB  two D2H streams %.1f GB/s total" % (mb, 2 * n / t / 1e9))
