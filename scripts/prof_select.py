import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
n, units = int(sys.argv[1]), int(sys.argv[2])
dev = torch.device("cuda:0")
_, _, _, std = bench.make_device_inputs(torch, n, units, 1, dev)
q = ops.q01_tensor([10.0 * (k % 101) / 100 for k in range(units)], dev)
for _ in range(5):
    thr = ops.select_threshold(std, units, q)
torch.cuda.synchronize()
print("ok", thr[:3])
