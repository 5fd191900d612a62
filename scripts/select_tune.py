"""Times pic_select_threshold on Kodak-shape units and checks the thresholds against torch.quantile.
Environment (read once per process by the library): PIC_TMA_SELECT, PIC_TMA_STAGES, PIC_TMA_CTAS, PIC_TMA_VPT."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
unit_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [101, 148, 296, 1010, 2048]
dev = torch.device("cuda:0")
_, _, _, std_all = bench.make_device_inputs(torch, n, max(unit_list), 1, dev)
L = pic_b200.lib()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
tag = " ".join(f"{k}={v}" for k, v in sorted(os.environ.items()) if k.startswith("PIC_"))
for units in unit_list:
    std = std_all[:units]
    prs = [10.0 * ((k * 7) % 101) / 100 for k in range(units)]
    q = ops.q01_tensor(prs, dev)
    thr = torch.empty(units, device=dev)
    ws = torch.empty(units * 4 + 256, dtype=torch.uint8, device=dev)
    fn = lambda: L.pic_select_threshold(std.data_ptr(), n, units, 0.5, q.data_ptr(), thr.data_ptr(), None, None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream)
    s0, f0 = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    L.pic_debug_select_counters(ctypes.byref(s0), ctypes.byref(f0))
    rc = fn()
    torch.cuda.synchronize()
    s1, f1 = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    L.pic_debug_select_counters(ctypes.byref(s1), ctypes.byref(f1))
    bad = 0
    for u in range(0, units, max(1, units // 64)):
        qq = float(q[u])
        if 0.0 <= qq <= 1.0:
            ref = torch.quantile(std[u], qq)
            if not (ref == thr[u] or (ref != ref and thr[u] != thr[u])):
                bad += 1
    t = timeit(fn)
    print(f"[{tag}] rc={rc} units={units:5d} {t:8.1f} us {units*n*4/t/1e3:7.0f} GB/s  sampled+{s1.value-s0.value} fallback+{f1.value-f0.value} mismatches={bad}", flush=True)
