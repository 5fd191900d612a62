"""Bandwidth of the un-fused drop-in kernels (GaussianConditional.forward / build_indexes / quantize) on one GPU.
usage: python scripts/elementwise_bench.py [elements]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pic_b200 as pic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 26
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
y = torch.randn(n, device=dev, generator=g) * 3
mu = torch.randn(n, device=dev, generator=g)
std = torch.rand(n, device=dev, generator=g) * 4 + 0.05
table = torch.exp(torch.linspace(torch.log(torch.tensor(0.11)), torch.log(torch.tensor(256.0)), 64)).to(dev)
sym = pic.ops.quantize(y, "symbols", means=mu)


def timed(fn, bytes_per_elem, name, reps=20):
    for _ in range(3):
        fn()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(reps):
        fn()
    ev[1].record()
    torch.cuda.synchronize()
    ms = ev[0].elapsed_time(ev[1]) / reps
    print(f"{name:28s} {ms * 1e3:8.1f} us  {n * bytes_per_elem / ms / 1e6:7.0f} GB/s (alloc included)")


timed(lambda: pic.ops.gaussian_forward(y, std, mu), 20, "gaussian_forward eval")
timed(lambda: pic.ops.gaussian_forward(y, std, mu, likelihood_only=True), 16, "gaussian likelihood only")
timed(lambda: pic.ops.build_indexes(std, table), 8, "build_indexes")
timed(lambda: pic.ops.quantize(y, "symbols", means=mu), 12, "quantize symbols")
timed(lambda: pic.ops.quantize(y, "dequantize", means=mu), 12, "quantize dequantize")
timed(lambda: pic.ops.dequantize(sym, mu), 12, "dequantize")
