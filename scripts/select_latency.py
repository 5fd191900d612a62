import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
dev = torch.device("cuda:0")
_, _, _, std_all = bench.make_device_inputs(torch, n, 4096, 1, dev)
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for units in (1, 37, 148, 296, 592, 888, 1010, 1776, 2048, 4096):
    std = std_all[:units]
    q = ops.q01_tensor([10.0 * ((k * 7) % 101) / 100 for k in range(units)], dev).clamp_(0.01, 0.99)
    thr = torch.empty(units, device=dev)
    ws = torch.empty(units * 4 + 256, dtype=torch.uint8, device=dev)
    L = pic_b200.lib()
    fn = lambda: L.pic_select_threshold(std.data_ptr(), n, units, 0.5, q.data_ptr(), thr.data_ptr(), None, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    t = timeit(fn)
    print(f"units={units:5d}  {t:8.1f} us   {units*n/t/1e3:7.1f} Gelem/s   {units*n*4/t/1e3:7.0f} GB/s")
