"""Dev helper: backward variants under SUSTAINED load (fwd + bwd back to back for ~1 s each, B = 256): what the bench sees once the
board sits at its power cap.  Reports ms per backward and the SM clock."""
import os, sys
import torch
sys.path.insert(0, ".")
from fireflies_b200.graphics import rasterization as R
import bench
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N, ts = 4096, [2048, 2048]
gen = torch.Generator().manual_seed(0)
pts = (torch.rand(N, 2, generator=gen) * 0.96 + 0.02).cuda()
ptsB = pts.unsqueeze(0).repeat(B, 1, 1).contiguous()
plan = R._SplatPlan(ptsB, B, 100.0, ts[0], ts[1], 4, 5)
gS = torch.randn(B, ts[0], ts[1], device="cuda"); gO = torch.randn(B, ts[1], ts[0], device="cuda")
S_, O_ = plan.forward(ptsB, True, True, True)
ev = lambda: torch.cuda.Event(enable_timing=True)
variants = [("st_rebuild", dict(FFB_SPLAT_BWD_PERSIST="1")), ("st_oneshot", dict()), ("st_saved", dict(FFB_SPLAT_BWD_SAVED="1")),
            ("old_saved", dict(FFB_SPLAT_BWD_ST="0")), ("st_rebuild", dict(FFB_SPLAT_BWD_PERSIST="1"))]
for name, kw in variants:
    for k in ("FFB_SPLAT_BWD_PERSIST", "FFB_SPLAT_BWD_SAVED", "FFB_SPLAT_BWD_ST"): os.environ.pop(k, None)
    os.environ.update(kw)
    for _ in range(3):
        plan.forward(ptsB, True, True, True); plan.backward(ptsB, gS, gO, True, O_)
    torch.cuda.synchronize()
    cs = bench.ClockSampler(0); cs.start()
    n = 60
    tf = tb = 0.0
    evs = []
    for i in range(n):
        a, b, c = ev(), ev(), ev()
        a.record(); plan.forward(ptsB, True, True, True); b.record(); plan.backward(ptsB, gS, gO, True, O_); c.record()
        evs.append((a, b, c))
    torch.cuda.synchronize()
    ck = cs.stop()
    half = evs[n // 2:]
    tf = sum(a.elapsed_time(b) for a, b, c in half) / len(half); tb = sum(b.elapsed_time(c) for a, b, c in half) / len(half)
    print(f"{name:12s}: fwd {tf:.3f} ms  bwd {tb:.3f} ms  (frac {B*8*ts[0]*ts[1]/tb/1e6/6458.4:.3f})  sm {ck['sm_mhz']} MHz {ck['reasons']}")
