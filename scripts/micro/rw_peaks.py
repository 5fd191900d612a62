"""Read-only, write-only and copy bandwidth of the GPU with stock torch kernels (context for the rooflines: the measured
peak of MEASURED_PEAKS.json is a COPY, i.e. half reads and half writes)."""
import torch
dev = torch.device("cuda:0")
n = 1 << 29                      # 2 GiB of f32
x = torch.randn(n, device=dev)
y = torch.empty_like(x)
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
gb = n * 4 / 1e9
print(f"read-only   x.sum()      {gb / t(lambda: x.sum()) * 1e3:8.0f} GB/s")
print(f"read-only   x.max()      {gb / t(lambda: x.max()) * 1e3:8.0f} GB/s")
print(f"write-only  y.zero_()    {gb / t(lambda: y.zero_()) * 1e3:8.0f} GB/s")
print(f"write-only  y.fill_(1.)  {gb / t(lambda: y.fill_(1.0)) * 1e3:8.0f} GB/s")
print(f"copy        y.copy_(x)   {2 * gb / t(lambda: y.copy_(x)) * 1e3:8.0f} GB/s (read + write bytes)")
