"""Probe: host-side cost of the single-sample API calls (configs[0]-style step)."""
import sys, time, torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
from fireflies_b200.graphics import rasterization as R
class P(dict):
    def update(self, *a, **k):
        return super().update(*a, **k) if (a or k) else None
gen = torch.Generator().manual_seed(0)
p0 = (torch.rand(100, 2, generator=gen) * 0.8 + 0.1).cuda().requires_grad_(True)
sc1 = ff.Scene(P())
m1 = ff.entity.Mesh("mesh-One", (torch.rand(10000, 3, generator=gen) * 2 - 1).cuda()); m1.rotate_z(-3.14159, 3.14159)
sc1._meshes.append(m1); sc1.train()
def timeit(name, fn, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name:34s} host {1e6 * (t1 - t0) / n:7.1f} us   incl. drain {1e6 * (t2 - t0) / n:7.1f} us")
timeit("mesh.randomize()", lambda: m1.randomize())
timeit("mesh.get_randomized_vertices()", lambda: m1.get_randomized_vertices())
timeit("splat_reduce fwd (no grad)", lambda: R.splat_reduce(p0.detach(), 100.0, [512, 512], sum_transposed=True))
def fb():
    p0.grad = None
    s, o = R.splat_reduce(p0, 100.0, [512, 512], sum_transposed=True)
    R.l1_loss(o, s).backward()
timeit("splat_reduce + l1 + backward", fb)
def fwd_only():
    s, o = R.splat_reduce(p0, 100.0, [512, 512], sum_transposed=True)
    return R.l1_loss(o, s)
timeit("splat_reduce + l1 (graph built)", fwd_only)
x = torch.rand(512, 512, device="cuda")
timeit("torch.empty x3 + tiny add", lambda: (torch.empty(512, 512, device="cuda"), torch.empty(512, 512, device="cuda"), x + 1))
