"""Probe: forward time per output combination at the config-3 shape (which output layout costs what)."""
import sys, torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, ts = 4096, [2048, 2048]
gen = torch.Generator().manual_seed(0)
pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
hw = ts[0] * ts[1]
for name, (ws, wo, st), nb in [("softor only", (False, True, False), 4), ("sum natural only", (True, False, False), 4),
                               ("sum transposed only", (True, False, True), 4), ("sum natural + softor", (True, True, False), 8),
                               ("sum transposed + softor", (True, True, True), 8)]:
    ms = t(lambda: plan.forward(ptsB, ws, wo, st))
    print(f"{name:26s} {ms:.3f} ms  {B*nb*hw/ms/1e6:.0f} GB/s")
