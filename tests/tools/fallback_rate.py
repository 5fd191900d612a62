"""Fallback rate / speed of the sampled select on spatially correlated std maps (real latents are smooth)."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch, torch.nn.functional as F
import pic_b200
from pic_b200 import ops
import pic_oracle as po
dev = torch.device("cuda:0")
L = pic_b200.lib()
def counters():
    a, b = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
    L.pic_debug_select_counters(ctypes.byref(a), ctypes.byref(b)); return a.value, b.value
g = torch.Generator(device=dev).manual_seed(3)
units, C, h, w = 512, 32, 32, 48
def field(kind):
    if kind == "iid":
        return torch.exp(torch.randn(units, C, h, w, device=dev, generator=g) * 1.2 - 1)
    if kind == "smooth8":   # 8x bilinear-upsampled low-res field: strong spatial correlation
        low = torch.randn(units, C, h // 8, w // 8, device=dev, generator=g)
        return torch.exp(F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) * 1.2 - 1)
    if kind == "smooth8+noise":
        low = torch.randn(units, C, h // 8, w // 8, device=dev, generator=g)
        return torch.exp(F.interpolate(low, size=(h, w), mode="bilinear", align_corners=False) * 1.2 - 1 + 0.05 * torch.randn(units, C, h, w, device=dev, generator=g))
    if kind == "channel-const":   # each channel nearly constant (worst case for row sampling)
        base = torch.randn(units, C, 1, 1, device=dev, generator=g)
        return torch.exp(base * 1.2 - 1 + 0.01 * torch.randn(units, C, h, w, device=dev, generator=g))
    if kind == "quantised":       # heavy ties
        return torch.round(torch.exp(torch.randn(units, C, h, w, device=dev, generator=g) * 1.2 - 1) * 16) / 16
for kind in ("iid", "smooth8", "smooth8+noise", "channel-const", "quantised"):
    std = field(kind).reshape(units, -1).contiguous()
    prs = [10.0 * ((k * 7) % 99 + 1) / 100 for k in range(units)]
    q = ops.q01_tensor(prs, dev)
    s0, f0 = counters()
    thr = ops.select_threshold(std, units, q)
    torch.cuda.synchronize()
    s1, f1 = counters()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.select_threshold(std, units, q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    # exactness on a subset against the oracle
    sub = list(range(0, units, 37))
    _, rthr = po.channel_mask(std[sub].cpu().numpy(), [prs[i] for i in sub])
    ok = np.array_equal(thr[sub].cpu().numpy(), rthr)
    print(f"{kind:15s} sampled={s1-s0:4d} fallback={f1-f0:4d}  select {ms*1e3:7.1f} us ({units*std.shape[1]/ms/1e6:6.1f} Gelem/s)  exact={ok}")
