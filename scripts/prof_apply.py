import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
n, units = int(sys.argv[1]), int(sys.argv[2])
mode = sys.argv[3] if len(sys.argv) > 3 else "apply"
dev = torch.device("cuda:0")
y_top, y_base, mu, std = bench.make_device_inputs(torch, n, units, 1, dev)
q = ops.q01_tensor([10.0 * (k % 101) / 100 for k in range(units)], dev)
table = pic_b200.get_scale_table().to(dev)
want = ("mask", "y_hat", "lik", "idx")
outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev) for k in want}
thr = ops.select_threshold(std, units, q)
for _ in range(5):
    ops.slice_forward(y_top, y_base, mu, std, units, q, table, thr_in=thr if mode == "apply" else None, want=want, out=outs)
torch.cuda.synchronize()
print("ok")
