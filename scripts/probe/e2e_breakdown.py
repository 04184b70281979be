"""Probe: where the end-to-end step (PatternStep.step_host) spends its time beyond the device-timed step."""
import sys, time, torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
from fireflies_b200.graphics import rasterization as R
from bench import build_scene, N_POINTS, TS, SIGMA
dev = torch.device("cuda", 0)
B = 256
pattern = (torch.rand(N_POINTS, 2, generator=torch.Generator().manual_seed(0)) * 0.96 + 0.02)
pts_host = pattern.pin_memory(); out_host = torch.empty(N_POINTS, 2).pin_memory(); loss_host = torch.empty(B).pin_memory()
sb = build_scene(ff, dev).batch(seed=1)
step = ff.PatternStep(N_POINTS, TS, SIGMA, B, scene_batch=sb, device=dev)
for i in range(3): step.step_host(pts_host, out_host, loss_host, sample0=i * B)
torch.cuda.synchronize()
K = 10
t0 = time.perf_counter()
for i in range(K): step.step_host(pts_host, out_host, loss_host, sample0=(3 + i) * B)
t1 = time.perf_counter()
print(f"step_host: {(t1 - t0) / K * 1e3:.3f} ms per step")
# launch-only time (no sync) and device time of the same call
ev = [torch.cuda.Event(True) for _ in range(2)]
torch.cuda.synchronize()
ev[0].record(); h0 = time.perf_counter()
for i in range(K): step.forward_backward(pts_host, sample0=i * B)
h1 = time.perf_counter(); ev[1].record(); torch.cuda.synchronize()
print(f"forward_backward back to back: host launch {(h1 - h0) / K * 1e3:.3f} ms, device {ev[0].elapsed_time(ev[1]) / K:.3f} ms per step")
# pieces
pts = step.pts_dev
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
plan = R._SplatPlan(pts, B, SIGMA, TS[0], TS[1], 4, 5)
s, o = plan.forward(pts, True, True, True)
print(f"plan {t(lambda: R._SplatPlan(pts, B, SIGMA, TS[0], TS[1], 4, 5)):.3f}  fwd {t(lambda: plan.forward(pts, True, True, True)):.3f}  "
      f"bwd_l1 {t(lambda: plan.backward_l1(pts, s, o, True)):.3f}  randomize {t(lambda: sb.randomize(B, sample0=0)):.3f}  "
      f"copy {t(lambda: step.pts_dev.copy_(pts_host.unsqueeze(0).expand_as(step.pts_dev), non_blocking=True)):.3f} ms")
