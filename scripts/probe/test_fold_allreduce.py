"""Multi-GPU check of the peer-memory fold + allreduce (torchrun, one rank per GPU): equality with fold + NCCL allreduce
over many epochs, the error flag, and the latency of both forms."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, ".")
local = int(os.environ["LOCAL_RANK"]); torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from fireflies_b200 import parallel as P
from fireflies_b200.graphics import rasterization as R
rank, world = dist.get_rank(), dist.get_world_size()
B, N = 64, 4096
f = P.FoldAllreduce(N * 2, dev)
ok = True
for it in range(40):
    x = torch.randn(B, N, 2, device=dev, generator=torch.Generator(device=dev).manual_seed(100 * it + rank))
    a = f(x)
    b = P.allreduce_sum_(R.reduce_over_samples(x))
    # same partial sums per rank; the cross-rank order differs (rank order vs NCCL's): compare with a tolerance
    ok &= bool(torch.allclose(a, b, rtol=1e-5, atol=1e-4))
    g = [torch.empty_like(a) for _ in range(world)]
    dist.all_gather(g, a)
    ok &= all(torch.equal(g[0], t) for t in g)          # bit-identical on every rank
f.check()
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
x = torch.randn(256, N, 2, device=dev)
t_f = t(lambda: f(x)); t_n = t(lambda: P.allreduce_sum_(R.reduce_over_samples(x)))
if rank == 0:
    print(f"world {world}: fused fold+allreduce {'OK' if ok else 'MISMATCH'}; fused {t_f:.1f} us, fold + NCCL {t_n:.1f} us per call")
dist.destroy_process_group()
