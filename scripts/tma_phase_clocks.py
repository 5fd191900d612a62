"""Role timers (clock64) of block 0 of the TMA select (debug build with -DPIC_PHASE_TIMING in build_variants/)."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
L = ctypes.CDLL(os.path.join(ROOT, "build_variants", "libpic_tma_dbg.so"))
vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
L.pic_select_threshold.argtypes = [vp, i64, i64, f32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 49152
qq = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
_, _, _, std_all = bench.make_device_inputs(torch, n, 2048, 1, dev)
names = ["sw:wait_final", "sw:wait_piv", "sw:sweep", "sw:wait_data(in sweep)", "h:pivots", "h:wait_sweep", "h:finish", "p:wait_empty", "f:count", "f:zero", "f:pass1", "f:find", "f:pass2", "f:rank+write"]
for units in (1, 148, 1010):
    std = std_all[:units]
    thr = torch.empty(units, device=dev)
    for _ in range(2):
        L.pic_select_threshold(std.data_ptr(), n, units, qq, None, thr.data_ptr(), None, None, None, 0, torch.cuda.current_stream().cuda_stream)
    clk = (ctypes.c_longlong * 16)()
    L.pic_debug_tma_phase_clocks(clk, 1)
    L.pic_select_threshold(std.data_ptr(), n, units, qq, None, thr.data_ptr(), None, None, None, 0, torch.cuda.current_stream().cuda_stream)
    L.pic_debug_tma_phase_clocks(clk, 0)
    per = (units + 147) // 148
    print(f"units={units:5d} ({per} per CTA) cycles per unit: " + "  ".join(f"{nm} {clk[i] // per:6d}" for i, nm in enumerate(names)))
