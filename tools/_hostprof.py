import cProfile, pstats, io, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ultrasphere_b200 as us
dev = torch.device("cuda:0")
c = us.create_standard(4)
k = torch.rand(5, device=dev, dtype=torch.float64)
f = lambda s: torch.exp(c.to_cartesian_dot(s, k))
run = lambda: us.integrate(c, f, False, 16, xp=torch, device=dev, dtype=torch.float64)
for _ in range(20): run()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): run()
torch.cuda.synchronize()
print("us per integral (n=16, fused):", (time.perf_counter() - t0) / 200 * 1e6)
pr = cProfile.Profile(); pr.enable()
for _ in range(200): run()
torch.cuda.synchronize(); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28); print(s.getvalue()[:6000])
