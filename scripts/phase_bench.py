#!/usr/bin/env python
"""Times the phases of the fused path separately: select only, apply only (thresholds given), full."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import pic_b200
from pic_b200 import ops
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
units = int(sys.argv[2]) if len(sys.argv) > 2 else 1010
dev = torch.device("cuda:0")
y_top, y_base, mu, std = bench.make_device_inputs(torch, n, units, 1, dev)
prs = [10.0 * (k % 101) / 100 for k in range(units)]
q = ops.q01_tensor(prs, dev)
table = pic_b200.get_scale_table().to(dev)
want = ("mask", "y_hat", "lik", "idx")
outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev) for k in want}
thr = ops.select_threshold(std, units, q)

def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

el = units * n
t_sel = timeit(lambda: ops.select_threshold(std, units, q))
t_app = timeit(lambda: ops.slice_forward(y_top, y_base, mu, std, units, q, table, thr_in=thr, want=want, out=outs))
t_full = timeit(lambda: ops.slice_forward(y_top, y_base, mu, std, units, q, table, want=want, out=outs))
t_mask = timeit(lambda: ops.channel_mask(std, units, q))
print(f"n={n} units={units} elems={el/1e6:.1f}M")
print(f"select only : {t_sel*1e3:8.1f} us  {el/t_sel/1e6:7.1f} Gelem/s  ({el*4/t_sel/1e6:.0f} GB/s of std)")
print(f"apply only  : {t_app*1e3:8.1f} us  {el/t_app/1e6:7.1f} Gelem/s  ({el*32/t_app/1e6:.0f} GB/s algorithmic)")
print(f"full fused  : {t_full*1e3:8.1f} us  {el/t_full/1e6:7.1f} Gelem/s  ({el*32/t_full/1e6:.0f} GB/s algorithmic)")
print(f"mask only   : {t_mask*1e3:8.1f} us  {el/t_mask/1e6:7.1f} Gelem/s")
# torch reference points: copy bandwidth on this device
a = torch.empty(256 << 20, dtype=torch.float32, device=dev); b = torch.empty_like(a)
t_cp = timeit(lambda: b.copy_(a), 10)
print(f"torch copy 1 GiB: {t_cp*1e3:.1f} us -> {2*a.numel()*4/t_cp/1e6:.0f} GB/s")
import ctypes
a_, b_ = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
pic_b200.lib().pic_debug_select_counters(ctypes.byref(a_), ctypes.byref(b_))
print("sampled units:", a_.value, "fallback units:", b_.value)
