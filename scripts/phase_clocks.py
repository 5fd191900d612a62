"""clock64() phase breakdown of the sampled select (debug build with -DPIC_PHASE_TIMING, built on the fly)."""
import ctypes, os, subprocess, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
SO = os.path.join(ROOT, "gpurun_out", "libpic_phase_dbg.so")
if not os.path.isfile(SO):
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    csrc = os.path.join(ROOT, "efficient-pic-with-variance-aware-masking_b200", "csrc")
    subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-shared", "-Xcompiler", "-fPIC",
                    "-DPIC_PHASE_TIMING", "-DPIC_PHASE_BLOCK=0", "-o", SO, os.path.join(csrc, "pic_latent.cu"),
                    os.path.join(csrc, "pic_host.cu"), os.path.join(csrc, "pic_rans.cpp")], check=True)
L = ctypes.CDLL(SO)
vp, i64, f32 = ctypes.c_void_p, ctypes.c_int64, ctypes.c_float
L.pic_select_threshold.argtypes = [vp, i64, i64, f32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
dev = torch.device("cuda:0")
n = 49152
_, _, _, std_all = bench.make_device_inputs(torch, n, 2048, 1, dev)
for units in (1, 101, 148, 888):
    std = std_all[:units]
    thr = torch.empty(units, device=dev)
    for _ in range(3):
        L.pic_select_threshold(std.data_ptr(), n, units, 0.5, None, thr.data_ptr(), None, None, None, 0, torch.cuda.current_stream().cuda_stream)
    clk = (ctypes.c_longlong * 16)()
    L.pic_debug_phase_clocks(clk)
    c = list(clk)[:5]
    d = [(c[i + 1] - c[i]) for i in range(4)]
    print(f"units={units:5d} cycles: sample {d[0]:7d}  pivots {d[1]:7d}  sweep {d[2]:7d}  final {d[3]:7d}  total {c[4]-c[0]:7d}  (~{(c[4]-c[0])/1.965e3:.1f} us)")
