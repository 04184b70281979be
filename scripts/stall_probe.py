"""Dev helper: per-step phase times of the bench loop (no synchronize between steps), with and without the NVML clock sampler thread:
where do the occasional 4 ms stalls at the start of a step come from?"""
import os, sys
import torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
import bench
B, K = 256, int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda", 0)
g0 = torch.Generator().manual_seed(0)
pattern = (torch.rand(4096, 2, generator=g0) * 0.96 + 0.02).to(dev)
gS = torch.randn(B, 2048, 2048, device=dev); gO = torch.randn(B, 2048, 2048, device=dev)
scene = bench.build_scene(ff, dev); sb = scene.batch(seed=1)
step = ff.PatternStep(4096, (2048, 2048), 100.0, B, scene_batch=sb, device=dev)
for i in range(3):
    step.forward_backward(pattern, upstream=(gS, gO), sample0=i * B)
torch.cuda.synchronize()
for sampler in (False, True, False, True):
    clk = bench.ClockSampler(0) if sampler else None
    if clk: clk.start()
    torch.cuda.synchronize()
    marks = [dict() for _ in range(K)]
    for i in range(K):
        step.forward_backward(pattern, upstream=(gS, gO), sample0=i * B, marks=marks[i])
    torch.cuda.synchronize()
    if clk: print(clk.stop())
    tot = marks[0]["start"].elapsed_time(marks[-1]["end"]) / K
    rows = [(m["start"].elapsed_time(m["prepare"]), m["randomize0"].elapsed_time(m["randomize1"]), m["prepare"].elapsed_time(m["fwd"]),
             m["fwd"].elapsed_time(m["bwd"]), m["bwd"].elapsed_time(m["end"])) for m in marks]
    gaps = [marks[i]["end"].elapsed_time(marks[i + 1]["start"]) for i in range(K - 1)]
    print(f"sampler={sampler}: {tot:.3f} ms/step; prepare per step:", " ".join(f"{r[0]:.2f}" for r in rows))
    print("   bwd:", " ".join(f"{r[3]:.2f}" for r in rows))
    print("   gaps between steps:", " ".join(f"{g:.2f}" for g in gaps))
