"""One-rank NCCL communicator, one band of a tiled unit: event timing of the two tiled select protocols (and the input
of an ncu launch list: `ncu --metrics gpu__time_duration.sum ... python scripts/tiled_profile.py`)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, pic_b200, bench
from pic_b200 import ops
L = pic_b200.lib()
dev = torch.device("cuda:0")
n_local = int(sys.argv[1]) if len(sys.argv) > 1 else 4194304
units = 10
ident = (ctypes.c_ubyte * 128)()
comm = ctypes.c_void_p()
assert L.pic_dist_unique_id(ident) == 0 and L.pic_dist_comm_init(ident, 0, 1, ctypes.byref(comm)) == 0
_, _, _, std = bench.make_device_inputs(torch, n_local, units, 7, dev)
q = ops.q01_tensor([1.0] * units, dev)
thr = torch.empty(units, device=dev)
ws = torch.empty(int(L.pic_tiled_sampled_workspace_bytes(n_local, n_local, units, 1)), dtype=torch.uint8, device=dev)
fb = ctypes.c_int(0)
st = torch.cuda.current_stream().cuda_stream
def sampled():
    assert L.pic_tiled_select_threshold_sampled(std.data_ptr(), n_local, n_local, units, 0.5, q.data_ptr(), thr.data_ptr(), ws.data_ptr(), ws.numel(), comm, st, ctypes.byref(fb)) == 0
def rounds():
    assert L.pic_tiled_select_threshold(std.data_ptr(), n_local, n_local, units, 0.5, q.data_ptr(), thr.data_ptr(), ws.data_ptr(), ws.numel(), comm, st) == 0
p2p = ctypes.c_void_p()
a, b = ctypes.c_size_t(0), ctypes.c_size_t(0)
assert L.pic_dist_p2p_region_bytes(n_local, units, 1, ctypes.byref(a), ctypes.byref(b)) == 0
assert L.pic_dist_p2p_init(comm, 0, a.value, b.value, ctypes.byref(p2p)) == 0
status = torch.zeros(1, dtype=torch.int32, device=dev)
def peer():
    assert L.pic_tiled_select_threshold_p2p(std.data_ptr(), n_local, n_local, units, 0.5, q.data_ptr(), thr.data_ptr(), ws.data_ptr(), ws.numel(), p2p, status.data_ptr(), st) == 0
want = ops.select_threshold(std, units, q)
def graphed(fn):
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        global st
        st = s.cuda_stream
        fn(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            fn()
    st = torch.cuda.current_stream().cuda_stream
    return g.replay
for name, fn in (("p2p", peer), ("p2p+graph", graphed(peer)), ("rounds+graph", graphed(rounds)), ("sampled", sampled), ("rounds", rounds)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    assert torch.equal(thr, want), name
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us per select of {units} units x {n_local} local elements (fallbacks {fb.value})")
assert status.item() == 0
L.pic_dist_p2p_destroy(p2p)
L.pic_dist_comm_destroy(comm)
