"""Experiment: select + apply chunk by chunk on ONE stream, so that the apply's second read of std hits L2 (a chunk's std
is chunk_units x 192 KB; L2 is 126 MB).  parts = 1 is the normal step."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
dev = torch.device("cuda:0")
wl = bench.WORKLOADS["kodak_sweep"]
n, units = wl["n"], 1010
y_top, y_base, mu, std = bench.make_device_inputs(torch, n, units, 3234, dev)
q = torch.cat([ops.q01_tensor(wl["prs"], dev)] * 10).contiguous()
table = pic_b200.get_scale_table().to(dev)
want = ("mask", "y_hat", "lik", "idx")
outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev) for k in want}
thr = torch.empty(units, device=dev)
L = pic_b200.lib()

def step(parts):
    s = torch.cuda.current_stream().cuda_stream
    cuts = [units * i // parts for i in range(parts + 1)]
    for i in range(parts):
        lo, hi = cuts[i], cuts[i + 1]
        assert L.pic_select_threshold(std[lo:hi].data_ptr(), n, hi - lo, 0.5, q[lo:hi].data_ptr(), thr[lo:hi].data_ptr(), None, None, None, 0, s) == 0
        ops.slice_forward(y_top[lo:hi], y_base[lo:hi], mu[lo:hi], std[lo:hi], hi - lo, q[lo:hi], table, thr_in=thr[lo:hi], want=want,
                          out={k: v[lo:hi] for k, v in outs.items()})

for parts in (1, 2, 3, 4, 5, 7, 10):
    for _ in range(3): step(parts)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        step(parts)
    for _ in range(3): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): g.replay()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 50
    print(f"parts={parts}: {ms:.4f} ms/step  {units * n / ms / 1e6:.1f} Gelem/s", flush=True)
