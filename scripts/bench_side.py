"""Side measurements for the non-headline BASELINE configs (not the bench.py contract): post-processing throughput at
configs[3] (64 x 1024^2 frames) and batched scene randomisation at configs[1] (vocal-fold scene, B=32).
CUDA events, inputs resident, L2 flushed between iterations for the small workloads."""
import json, sys
import torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
from fireflies_b200.postprocessing.base import run_postprocess

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if __import__("os").path.isfile("MEASURED_PEAKS.json") else 6650.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, n=10, do_flush=True):
    for _ in range(3): fn()
    ev = [(torch.cuda.Event(True), torch.cuda.Event(True)) for _ in range(n)]
    for a, b in ev:
        if do_flush: flush.zero_()
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in ev)
    return t[len(t) // 2]

out = {}
B, H, W = 64, 1024, 1024
x = torch.rand(B, H, W, device="cuda")
gates_on = torch.ones(B, 2, dtype=torch.uint8, device="cuda")
for name, kw in (("blur3x3+noise+clip", dict(blur=((3, 3), (5.0, 5.0)), noise=(0.0, 0.05))),
                 ("blur5x5", dict(blur=((5, 5), (3.0, 3.0)))), ("blur11x11", dict(blur=((11, 11), (5.0, 5.0)))),
                 ("noise+clip", dict(noise=(0.0, 0.05)))):
    ms = timed(lambda: run_postprocess(x, gates=gates_on, seed=1, frame0=0, **kw))
    gbs = 8 * B * H * W / (ms * 1e-3) / 1e9
    out["post_" + name] = {"ms_per_64_frames": ms, "frames_per_s": B / (ms * 1e-3), "GBs": gbs, "frac_of_measured_peak": gbs / PEAK}

class P(dict):
    def update(self, *a, **k):
        return super().update(*a, **k) if (a or k) else None
g = torch.Generator().manual_seed(2)
F, V1, V2, Bs = 64, 20000, 50000, 32
frames = (torch.rand(F, V1, 3, generator=g) * 2 - 1).cuda()
sc = ff.Scene(P())
vf = ff.entity.Mesh("mesh-VocalFold", frames[0]); vf.add_train_animation(frames); vf.add_eval_animation(frames, max=F - 1)
vf.scale_x(0.5, 2.0); vf.rotate_y(-0.25, 0.25)
la = ff.entity.Mesh("mesh-Larynx", (torch.rand(V2, 3, generator=g) * 2 - 1).cuda()); la.scale_x(0.8, 1.2); la.rotate_y(-0.1, 0.1)
sc._meshes += [vf, la]; sc.train()
sb = sc.batch(seed=1)
ms = timed(lambda: sb.randomize(Bs))
out["scene_randomize_config1_B32"] = {"ms": ms, "samples_per_s": Bs / (ms * 1e-3), "GBs": 24 * (V1 + V2) * Bs / (ms * 1e-3) / 1e9}
# configs[0]: examples/09-style single pattern, 100 points into 512^2, L1(softor, sum) step through the autograd API + one rotate_z mesh
import time
from fireflies_b200.graphics import rasterization as R
gen = torch.Generator().manual_seed(0)
p0 = (torch.rand(100, 2, generator=gen) * 0.8 + 0.1).cuda().requires_grad_(True)
sc1 = ff.Scene(P())
m1 = ff.entity.Mesh("mesh-One", (torch.rand(10000, 3, generator=gen) * 2 - 1).cuda()); m1.rotate_z(-3.14159, 3.14159)
sc1._meshes.append(m1); sc1.train()
def c0_step():
    p0.grad = None
    m1.randomize(); m1.get_randomized_vertices()
    s, o = R.splat_reduce(p0, 100.0, [512, 512], sum_transposed=True)
    R.l1_loss(o, s).backward()
for _ in range(5): c0_step()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(50): c0_step()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 50
out["config0_single_sample_step"] = {"ms": dt * 1e3, "samples_per_s": 1.0 / dt,
                                     "note": "host-launch bound: randomize + vertices + bin + fwd + L1 + bwd through the reference-shaped API, wall clock"}
print(json.dumps(out, indent=1))
