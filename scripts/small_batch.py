import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
dev = torch.device("cuda:0")
table = pic_b200.get_scale_table().to(dev)
want = ("mask", "y_hat", "lik", "idx")
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n in (8192, 49152, 131072):
    for units in (1, 4, 16, 32, 101, 256):
        y_top, y_base, mu, std = bench.make_device_inputs(torch, n, units, 1, dev)
        outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev) for k in want}
        t = timeit(lambda: ops.slice_forward(y_top, y_base, mu, std, units, 0.75, table, want=want, out=outs))
        print(f"n={n:6d} units={units:4d}  {t:8.1f} us  {units*n/t/1e3:7.1f} Gelem/s")
