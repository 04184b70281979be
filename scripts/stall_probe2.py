"""Dev helper: 60 bench-like timed regions of 10 steps; print every step whose first phase starts late, with its index in the region."""
import gc, os, sys, time
import torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
import bench
B, K, R = 256, 10, int(sys.argv[1]) if len(sys.argv) > 1 else 60
dev = torch.device("cuda", 0)
g0 = torch.Generator().manual_seed(0)
pattern = (torch.rand(4096, 2, generator=g0) * 0.96 + 0.02).to(dev)
gS = torch.randn(B, 2048, 2048, device=dev); gO = torch.randn(B, 2048, 2048, device=dev)
scene = bench.build_scene(ff, dev); sb = scene.batch(seed=1)
step = ff.PatternStep(4096, (2048, 2048), 100.0, B, scene_batch=sb, device=dev)
for i in range(3):
    step.forward_backward(pattern, upstream=(gS, gO), sample0=i * B)
torch.cuda.synchronize()
late = 0
for r in range(R):
    time.sleep(0.05 if len(sys.argv) > 2 else 0.0)          # optional idle gap between regions (the GPU drops its clocks when idle)
    torch.cuda.synchronize()
    marks = [dict() for _ in range(K)]
    host = []
    for i in range(K):
        t0 = time.perf_counter()
        step.forward_backward(pattern, upstream=(gS, gO), sample0=i * B, marks=marks[i])
        host.append((time.perf_counter() - t0) * 1e3)
    torch.cuda.synchronize()
    for i, m in enumerate(marks):
        p = m["start"].elapsed_time(m["prepare"])
        if p > 0.6:
            late += 1
            print(f"region {r} step {i}: prepare {p:.2f} ms, randomize {m['randomize0'].elapsed_time(m['randomize1']):.2f}, "
                  f"start->randomize0 {m['start'].elapsed_time(m['randomize0']):.2f}, host enqueue of that step {host[i]:.2f} ms (all: {' '.join(f'{h:.1f}' for h in host)})")
print("late steps:", late, "of", R * K)
