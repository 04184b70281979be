"""Dev helper: bench-like A/B of the backward launch forms inside the full step (3 warm-up + 10 timed steps per sample, alternating,
5 rounds): mean / min / max ms per step.  What the driver's bench sees, including the board's power management."""
import os, sys, statistics
import torch
sys.path.insert(0, ".")
import fireflies_b200 as ff
import bench
B = 256
dev = torch.device("cuda", 0)
g0 = torch.Generator().manual_seed(0)
pattern = (torch.rand(4096, 2, generator=g0) * 0.96 + 0.02).to(dev)
gS = torch.randn(B, 2048, 2048, device=dev); gO = torch.randn(B, 2048, 2048, device=dev)
scene = bench.build_scene(ff, dev); sb = scene.batch(seed=1)
step = ff.PatternStep(4096, (2048, 2048), 100.0, B, scene_batch=sb, device=dev)
variants = {"persist": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_GRID": str(148 * 24)}, "oneshot": {}, "persist_g20": {"FFB_SPLAT_BWD_PERSIST": "1"},
            "persist_g16": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_GRID": str(148 * 16)}, "old": {"FFB_SPLAT_BWD_ST": "0"},
            "chunk2": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_CHUNK": "2"}, "chunk4": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_CHUNK": "4"},
            "chunk8": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_CHUNK": "8"}, "chunk32": {"FFB_SPLAT_BWD_PERSIST": "1", "FFB_SPLAT_BWD_CHUNK": "32"},
            "l1_old": {"L1": "1", "FFB_SPLAT_L1_ST": "0"}, "l1_st": {"L1": "1"}, "l1_st_g24": {"L1": "1", "FFB_SPLAT_BWD_GRID": str(148 * 24)}}
if len(sys.argv) > 1: variants = {k: v for k, v in variants.items() if k in sys.argv[1:]}
res = {k: [] for k in variants}
ev = lambda: torch.cuda.Event(enable_timing=True)
for rnd in range(5):
    for name, env in variants.items():
        for k in ("FFB_SPLAT_BWD_PERSIST", "FFB_SPLAT_BWD_GRID", "FFB_SPLAT_BWD_ST", "FFB_SPLAT_L1_ST", "FFB_SPLAT_BWD_CHUNK"): os.environ.pop(k, None)
        os.environ.update({k: v for k, v in env.items() if k != "L1"})
        up = None if "L1" in env else (gS, gO)
        for i in range(3): step.forward_backward(pattern, upstream=up, sample0=i * B)
        torch.cuda.synchronize()
        e0, e1 = ev(), ev(); e0.record()
        for i in range(10): step.forward_backward(pattern, upstream=up, sample0=(3 + i) * B)
        e1.record(); torch.cuda.synchronize()
        res[name].append(e0.elapsed_time(e1) / 10)
for k, v in res.items():
    print(f"{k:12s} mean {statistics.mean(v):.3f}  min {min(v):.3f}  max {max(v):.3f} ms/step  -> {B / statistics.mean(v):.1f} k samples/s   runs: " + " ".join(f"{x:.2f}" for x in v))
