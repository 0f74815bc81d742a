"""HBM bandwidth probe: pure write (fill), pure read (sum), copy."""
import torch
dev = torch.device('cuda:0')
n = 1 << 30
a = torch.empty(n, dtype=torch.float32, device=dev)   # 4 GiB
b = torch.empty(n, dtype=torch.float32, device=dev)
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
gb = n * 4 / 1e9
print('fill  (write only): %.0f GB/s' % (gb / t(lambda: a.fill_(1.0))))
print('memset (zero_)    : %.0f GB/s' % (gb / t(lambda: a.zero_())))
print('sum   (read only) : %.0f GB/s' % (gb / t(lambda: a.sum())))
print('copy  (r + w)     : %.0f GB/s' % (2 * gb / t(lambda: b.copy_(a))))
print('add   (2r + w)    : %.0f GB/s' % (3 * gb / t(lambda: torch.add(a, b, out=a))))
