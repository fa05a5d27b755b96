"""Write-only and read-only HBM bandwidth next to the copy figure (torch fill / sum / copy over 4 GB)."""
import torch
n = 1 << 30
x = torch.empty(n, dtype=torch.float32, device="cuda"); y = torch.empty_like(x)
def t(f, reps=5):
    f(); torch.cuda.synchronize()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): f()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e-3
s = t(lambda: x.zero_()); print("write-only %.0f GB/s" % (4 * n / s / 1e9))
s = t(lambda: x.sum()); print("read-only %.0f GB/s" % (4 * n / s / 1e9))
s = t(lambda: y.copy_(x)); print("copy %.0f GB/s (read+write)" % (8 * n / s / 1e9))
