"""Probe: does the power-of-two row pitch (8 KB at 2048 columns) cost DRAM efficiency?  Same density, different sides."""
import sys, torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R
B = 64
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
for side in (2048, 2080, 2112, 1984, 2304):
    N = int(4096 * (side / 2048.0) ** 2)
    ts = [side, side]
    gen = torch.Generator().manual_seed(0)
    pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
    ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
    plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
    gS = torch.randn(B, ts[0], ts[1], device="cuda"); gO = torch.randn(B, ts[1], ts[0], device="cuda")
    tf = t(lambda: plan.forward(ptsB, True, True, True))
    S_, O_ = plan.forward(ptsB, True, True, True)
    tb = t(lambda: plan.backward(ptsB, gS, gO, True, O_))
    hw = side * side
    print(f"side {side}: fwd {tf:.3f} ms {B*8*hw/tf/1e6:.0f} GB/s | bwd {tb:.3f} ms actual {B*12*hw/tb/1e6:.0f} GB/s (algorithmic {B*8*hw/tb/1e6:.0f})")
    del gS, gO, S_, O_, plan
