"""Dev helper: the one point of test_every_kernel_variant (ns 3, no 2, soft-OR only) that sits at 1.08 tolerance units: who is right?"""
import os, sys
import torch
sys.path.insert(0, ".")
from oracle import ff_oracle as O
from fireflies_b200.graphics import rasterization as R
gen = torch.Generator().manual_seed(33)
B, N, ts, sigma = 2, 250, [288, 224], 36.0
pts = (torch.rand(B, N, 2, generator=gen) * 0.98 + 0.01)
wS = torch.randn(B, ts[1], ts[0], generator=gen)
wO = torch.randn(B, ts[1], ts[0], generator=gen)
ns, no = 3, 2
plan = R._SplatPlan(pts.cuda(), B, sigma, ts[0], ts[1], ns, no)
s, o = plan.forward(pts.cuda(), False, True, False)
g = plan.backward(pts.cuda(), None, wO.cuda(), False, None).cpu().double()
for b in range(B):
    ana = O.splat_grad_analytic(pts[b], sigma, ts, None, wO[b], ns, no)
    nrm = ana.norm(dim=1, keepdim=True)
    e = (g[b] - ana).abs() / (1e-4 * ana.abs() + 1e-4 * nrm + 1e-12)
    i = int(e.max(dim=1).values.argmax())
    print(f"sample {b}: worst point {i}: err {e[i].tolist()} kernel {g[b][i].tolist()} analytic {ana[i].tolist()} norm {float(nrm[i])}, mean norm {float(nrm.mean())}")
    P = pts[b].float() * torch.tensor([ts[0], ts[1]], dtype=torch.float32)
    print("   P", P[i].tolist(), "frac", (P[i] - P[i].floor()).tolist(), "P-half", (P[i] - 6).tolist())
    d = (P - P[i]).abs().max(dim=1).values
    nb = torch.nonzero(d < 14).flatten().tolist()
    print("   neighbours", [(j, P[j].tolist()) for j in nb if j != i])
    # forward check at this point's window
    of = O.baked_softor(pts[b], sigma, ts, no)
    print("   forward max |diff| whole texture", float((o[b].cpu() - of).abs().max()))
    # top error points
    top = e.max(dim=1).values.topk(5)
    print("   top5", [(int(k), round(float(v), 3)) for v, k in zip(top.values, top.indices)])
