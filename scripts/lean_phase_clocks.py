"""clock64 phase breakdown of block 0 of the lean select (debug build with -DPIC_PHASE_TIMING in build_variants/)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from pic_b200 import ops
L = ctypes.CDLL(os.path.join(ROOT, "build_variants", "libpic_tma_dbg.so"))
vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
L.pic_select_threshold.argtypes = [vp, i64, i64, f32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
_, _, _, std_all = bench.make_device_inputs(torch, n, 2048, 1, dev)
for units in (1, 148, 444, 1010, 2048):
    std = std_all[:units]
    q = ops.q01_tensor([10.0 * ((k * 7) % 101) / 100 for k in range(units)], dev).clamp_(0.01, 0.99)
    thr = torch.empty(units, device=dev)
    call = lambda: L.pic_select_threshold(std.data_ptr(), n, units, 0.5, q.data_ptr(), thr.data_ptr(), None, None, None, 0, torch.cuda.current_stream().cuda_stream)
    for _ in range(2): call()
    clk = (ctypes.c_longlong * 16)()
    L.pic_debug_tma_phase_clocks(clk, 1)
    call()
    L.pic_debug_tma_phase_clocks(clk, 0)
    grid = min(units, 444 if units > 148 else 148)
    per = (units + grid - 1) // grid
    print(f"units={units:5d} ({per} per CTA) cycles per unit of block 0: pivots {clk[0] // per:6d}  sweep {clk[1] // per:6d}  reduce+zero {clk[6] // per:6d}  pass1 {clk[3] // per:6d}  find {clk[4] // per:6d}  pass2 {clk[5] // per:6d}  rank+write+sync {clk[2] // per:6d}")
