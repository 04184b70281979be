"""Dev helper: time the fused splat fwd / bwd at BASELINE config-3 shape (not the bench contract)."""
import sys, time
import torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N, ts = 4096, [2048, 2048]
gen = torch.Generator().manual_seed(0)
pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
gS = torch.randn(B, ts[0], ts[1], device="cuda")
gO = torch.randn(B, ts[1], ts[0], device="cuda")
def t(fn, n=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
tp = t(lambda: R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5))
tf = t(lambda: plan.forward(ptsB, True, True, True))
S_, O_ = plan.forward(ptsB, True, True, True)
tb = t(lambda: plan.backward(ptsB, gS, gO, True, O_))
tb2 = t(lambda: plan.backward(ptsB, gS, gO, True))
import os
os.environ["FFB_SPLAT_NO_TMA"] = "1"
tf3 = t(lambda: plan.forward(ptsB, True, True, True))
S3, O3 = plan.forward(ptsB, True, True, True)
print(f"fwd without TMA: {tf3:.3f} ms; tma == plain: sum {bool(torch.equal(S3, S_))} softor {bool(torch.equal(O3, O_))}")
tb3 = t(lambda: plan.backward(ptsB, gS, gO, True, O_))
d_old = plan.backward(ptsB, gS, gO, True, O_)
os.environ["FFB_SPLAT_NO_TMA"] = "0"
d_new = plan.backward(ptsB, gS, gO, True, O_)
print(f"bwd (saved) without TMA: {tb3:.3f} ms; max |tma - plain| = {float((d_new - d_old).abs().max()):.3e} of {float(d_old.abs().max()):.3e}")
print(f"bwd without saved softor: {tb2:.3f} ms")
os.environ["FFB_SPLAT_EAGER"] = "0"
tb4 = t(lambda: plan.backward(ptsB, gS, gO, True, O_))
os.environ["FFB_SPLAT_EAGER"] = "1"
tb5 = t(lambda: plan.backward(ptsB, gS, gO, True, O_))
del os.environ["FFB_SPLAT_EAGER"]
print(f"bwd eager off {tb4:.3f} ms, on {tb5:.3f} ms")
tl = t(lambda: plan.backward_l1(ptsB, S_, O_, True))
print(f"fused L1 backward: {tl:.3f} ms ({B*16*ts[0]*ts[1]/tl/1e6:.0f} GB/s of 16 B/texel actual reads)")
hw = ts[0] * ts[1]
print(f"B={B} prepare {tp:.3f} ms  fwd {tf:.3f} ms ({B*8*hw/tf/1e6:.0f} GB/s)  bwd {tb:.3f} ms ({B*8*hw/tb/1e6:.0f} GB/s)"
      f"  fwd+bwd per sample {(tf+tb)/B*1e3:.2f} us -> {B/(tf+tb)*1e3:.0f} samples/s, roofline frac {(B*(16*hw+16*N)/((tp+tf+tb)*1e-3))/6445.6e9:.3f}")
