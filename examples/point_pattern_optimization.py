"""Laser-pattern regularisation on the B200 -- the loop of the reference's ``test_point_reg``
(fireflies/graphics/rasterization.py:564-643; BASELINE configs[0], ``examples/09_point_pattern_optimization.py`` is an empty
file in the reference): spread N laser points so that their splats overlap as little as possible by minimising
``L1Loss(softor(tex), sum(tex))`` with Adam.

Two forms of the same step:
  --api reference   the reference's own lines, unchanged: ``baked_sum_2`` / ``baked_softor_2`` / ``torch.nn.L1Loss`` (two
                    splat launches and torch's loss per step; autograd flows through the CUDA kernels);
  --api fused       ``splat_reduce`` (one prepare + one fused forward for both reductions) and ``l1_loss`` (the fused
                    L1 loss + backward kernel), the form ``PatternStep`` batches over many scenes.
Needs a CUDA device (there is no CPU path).  Usage: python examples/point_pattern_optimization.py [--api fused] [--steps 200]
"""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fireflies_b200 as fireflies  # noqa: E402

R = fireflies.graphics.rasterization


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--api", default="reference", choices=["reference", "fused"])
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--points", type=int, default=500)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--increase-overlap", action="store_true", help="maximise the overlap instead (reduce_overlap=False)")
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("this example needs a CUDA device")
    device = torch.device("cuda")
    torch.manual_seed(0)

    points = torch.rand([args.points, 2], device=device)
    points.requires_grad = True
    sigma = torch.tensor([15.0], device=device) ** 2
    texture_size = torch.tensor([args.size, args.size], device=device)
    loss_func = torch.nn.L1Loss()
    optim = torch.optim.Adam([{"params": points, "lr": 0.001}])
    sign = -1.0 if args.increase_overlap else 1.0

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.steps):
        optim.zero_grad()
        if args.api == "reference":
            summed = R.baked_sum_2(points, sigma, texture_size)
            softored = R.baked_softor_2(points, sigma, texture_size)
            loss = sign * loss_func(softored, summed)
        else:
            summed, softored = R.splat_reduce(points, float(sigma), [args.size, args.size], sum_transposed=True)
            loss = sign * R.l1_loss(softored, summed)
        loss.backward()
        optim.step()
        with torch.no_grad():
            points[points >= 1.0] = 0.999
            points[points <= 0.0] = 0.001
        if i % 20 == 0 or i == args.steps - 1:
            print(f"step {i:4d}  loss {loss.item():.6f}")
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"{args.steps} steps in {dt:.2f} s ({args.steps / dt:.0f} steps/s, api={args.api})")


if __name__ == "__main__":
    main()
