"""Dev helper: persistent backward, time vs grid size (FFB_SPLAT_BWD_GRID)."""
import os, sys
import torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R
B = 64
N, ts = 4096, [2048, 2048]
gen = torch.Generator().manual_seed(0)
pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
gS = torch.randn(B, ts[0], ts[1], device="cuda")
gO = torch.randn(B, ts[1], ts[0], device="cuda")
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for g in sys.argv[1:]:
    os.environ["FFB_SPLAT_BWD_GRID"] = g
    print(g, f"{t(lambda: plan.backward(ptsB, gS, gO, True)):.3f} ms")
