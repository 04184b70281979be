"""Probe: pure-write vs copy bandwidth on this GPU (what bounds a write-only kernel like the splat forward)."""
import torch
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    best = 1e9
    for _ in range(n):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device="cuda")
b = torch.empty(n, dtype=torch.float32, device="cuda")
print(f"fill_  : {n*4/t(lambda: a.fill_(1.0))/1e6:.0f} GB/s (write only)")
print(f"zero_  : {n*4/t(lambda: a.zero_())/1e6:.0f} GB/s (memset)")
print(f"copy_  : {2*n*4/t(lambda: b.copy_(a))/1e6:.0f} GB/s (read+write)")
print(f"sum    : {n*4/t(lambda: a.sum())/1e6:.0f} GB/s (read only)")
print(f"a+b->a : {3*n*4/t(lambda: torch.add(a, b, out=a))/1e6:.0f} GB/s (2 reads + 1 write)")
