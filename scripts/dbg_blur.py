import sys, torch
sys.path.insert(0, ".")
from fireflies_b200.postprocessing.base import run_postprocess
shape = tuple(int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (3, 64, 128)
x = torch.rand(shape, generator=torch.Generator().manual_seed(1)).cuda()
y = run_postprocess(x, blur=((3, 3), (5.0, 5.0)))
torch.cuda.synchronize()
print("ok", y.sum().item())
