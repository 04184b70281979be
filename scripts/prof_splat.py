"""Dev helper for ncu: config-3 splat, 3 iterations of exactly [prepare, fwd, bwd] (B from argv, default 64)."""
import sys
import torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, ts = 4096, [2048, 2048]
gen = torch.Generator().manual_seed(0)
pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
gS = torch.randn(B, ts[0], ts[1], device="cuda")
gO = torch.randn(B, ts[1], ts[0], device="cuda")
for _ in range(3):
    plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
    S, O = plan.forward(ptsB, True, True, True)
    d = plan.backward(ptsB, gS, gO, True, O)
    if len(sys.argv) > 2:
        l = plan.backward_l1(ptsB, S, O, True)
torch.cuda.synchronize()
print("ok", float(d.abs().sum()))
