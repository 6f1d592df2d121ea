import torch, time
n = 1 << 28  # 1 GiB of float32
h = torch.empty(n, dtype=torch.float32).pin_memory()
d = torch.empty(n, dtype=torch.float32, device="cuda")
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); best = min(best, time.perf_counter() - t0)
    return best
print("H2D %.1f GB/s" % (n * 4 / t(lambda: d.copy_(h, non_blocking=True)) / 1e9))
print("D2H %.1f GB/s" % (n * 4 / t(lambda: h.copy_(d, non_blocking=True)) / 1e9))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h2 = torch.empty(n, dtype=torch.float32).pin_memory(); d2 = torch.empty(n, dtype=torch.float32, device="cuda")
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
dt = t(both); print("bidirectional: %.1f GB/s each way" % (n * 4 / dt / 1e9))
