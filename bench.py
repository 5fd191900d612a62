#!/usr/bin/env python
"""bench.py -- mask+quantize+likelihood throughput of the PIC latent hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N>1: one rank per GPU

The printed line measures `kodak_sweep` (BASELINE config[1]); its "workloads" block carries one record per other
BASELINE configuration, measured in the same run (skip with --only-main):

  kodak_sweep (headline): Kodak-shape 768x512 latents ([1,32,32,48] => n=49152 per unit), 10 progressive slices x
      101-point quality sweep (pr = 0, 0.1, .. 10) = 1010 units per step with INDEPENDENT latents per unit, codec
      outputs (mask, y_hat, likelihood, scale index).  One step = select launch + apply launch over all units.
      Weak scaling: every rank runs a full replica on its own data, no collective on the data path.
  per_slice: the same units in encoder order -- one select + one apply launch per slice index (10 dependent slices).
  sweep_shared: the quality sweep as the reference runs it (check_levels_np): per slice ONE set of latents evaluated
      at the 101 qualities (pic_slice_forward_multi: inputs cross HBM once, 16 B/element of output).
  first_train (config[2]): [256,32,16,16] x 10 slices, random quality per image, training forward (noise) + backward.
  rem_latent (config[3]): REM variant at the first_train shape: 3 threshold selections per slice (pre-REM attention
      mask duplicated to [B,64,h,w], checkpoint pass at q = 0.75, post-REM block mask) + slice.
  tile8192 (config[4]): one 8192x8192 image, 10 slices of n=8388608; at N > 1 each rank holds a row band and the
      per-slice thresholds come from NCCL all-reduced radix histograms (strong scaling, a real collective).
Inputs of one step are far larger than the 126 MB L2 (kodak_sweep: 794 MB), so consecutive steps stream from HBM
without an explicit flush.  One JSON line is printed by rank 0 (keys: DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

WORKLOADS = {
    "kodak_sweep": dict(n=32 * 32 * 48, slices=10, prs=[10.0 * k / 100 for k in range(101)], scaling="weak",
                        desc="Kodak 768x512 batch 1, 10 progressive slices x 101-point quality sweep, codec outputs",
                        bytes_per_elem=32,  # y_top, y_base, mu, std read once (std 2nd read from L2) + mask, y_hat, lik, idx
                        ),
    "first_train": dict(n=32 * 16 * 16, slices=10, prs=None, batch=256, scaling="weak",
                        desc="256 crops of 256x256, 10 slices, random q per image, training forward + backward",
                        bytes_per_elem=32 + 48,  # fwd: 5 in (incl. noise) + 3 out; bwd: 8 in + 4 out
                        ),
    "rem_latent": dict(n=32 * 16 * 16, slices=10, prs=None, batch=256, scaling="weak",
                       desc="REM variant (mu_std): per slice 3 threshold selections (pre-REM star mask duplicated to "
                            "[B,64,h,w], checkpoint pass at q=0.75, post-REM block mask) + the fused slice forward",
                       bytes_per_elem=4 + 8 + 4 + 32,  # star select + duplicated mask write, checkpoint select, slice fwd
                       ),
    "tile8192": dict(n=32 * 512 * 512, slices=10, prs=[1.0], scaling="strong",
                     desc="single 8192x8192 synthetic image, 10 slices, q=1, row bands over the ranks, NCCL histogram all-reduce",
                     bytes_per_elem=36,  # std read by the select (>=1 pass from HBM) + the 32 of the apply
                     ),
}
METRIC = "mask+quantize+likelihood throughput"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """nvidia-smi samples of every GPU of the job during the timed region (rank 0 runs it)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int, n_gpus: int = 1):
        self.ids = [index] if n_gpus <= 1 else list(range(n_gpus))
        self.rows, self.proc = [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--id=" + ",".join(map(str, self.ids)), f"--query-gpu={self.FIELDS}",
                 "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        inside, anytime, smax, reasons = {}, {}, [], set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                gpu, clk, mx = int(parts[0]), float(parts[1]), float(parts[2])
            except ValueError:
                continue
            smax.append(mx)
            anytime.setdefault(gpu, []).append(clk)
            if t0 <= ts <= t1 + 0.03:
                inside.setdefault(gpu, []).append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        # a GPU whose line of the one sample inside a short timed region arrived late falls back to its other samples
        use = {g: (inside.get(g) or anytime.get(g)) for g in set(inside) | set(anytime)}
        med = {g: statistics.median(v) for g, v in use.items() if v}
        out = {"sm_mhz": min(med.values()) if med else None, "sm_max_mhz": max(smax) if smax else None,
               "reasons": sorted(reasons), "samples_in_timed_region": min((len(v) for v in inside.values()), default=0)}
        if len(self.ids) > 1:
            out["sm_mhz_by_gpu"] = [med.get(g) for g in self.ids]   # sm_mhz is the slowest GPU's median
        return out


# ------------------------------------------------------------------------------------------ inputs
def make_host_inputs(n, units, seed):
    """'S-trained-like' synthetic latents on the host (numpy)."""
    rng = np.random.default_rng(seed)
    shape = (units, n)
    std = np.exp(rng.normal(-1.0, 1.2, size=shape)).clip(1e-3, 300.0).astype(np.float32)
    std[rng.random(size=shape) < 0.02] *= -1
    mu = rng.normal(0, 1, size=shape).astype(np.float32)
    y_base = rng.normal(0, 2, size=shape).astype(np.float32)
    y_top = (y_base + mu + np.abs(std) * rng.normal(0, 1, size=shape).astype(np.float32)).astype(np.float32)
    return y_top, y_base, mu, std


def make_device_inputs(torch, n, units, seed, dev):
    g = torch.Generator(device=dev).manual_seed(seed)
    shape = (units, n)
    std = torch.exp(torch.randn(shape, device=dev, generator=g) * 1.2 - 1.0).clamp_(1e-3, 300.0)
    flip = torch.rand(shape, device=dev, generator=g) < 0.02
    std = torch.where(flip, -std, std)
    mu = torch.randn(shape, device=dev, generator=g)
    y_base = torch.randn(shape, device=dev, generator=g) * 2
    y_top = y_base + mu + std.abs() * torch.randn(shape, device=dev, generator=g)
    return y_top.contiguous(), y_base.contiguous(), mu.contiguous(), std.contiguous()


def unit_prs(wl, units_per_slice, seed):
    if wl["prs"] is not None:
        return list(wl["prs"])
    rng = np.random.default_rng(seed)
    return [float(v) for v in rng.uniform(0.05, 9.95, size=units_per_slice)]


# ------------------------------------------------------------------------------------------ reference arm
def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1; ignore that)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_step_units(wl):
    """Units of one CPU step: the WHOLE step of the CUDA arm (same units, same qualities), except for the
    8M-element units of tile8192 (one unit is ~1 s of qsort)."""
    n = wl["n"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    total = per_slice * wl["slices"]
    prs_slice = unit_prs(wl, per_slice, 4321)
    if n > (1 << 20):
        return 1, prs_slice[:1], total
    return total, [prs_slice[u % per_slice] for u in range(total)], total


def cpu_pass(po, arrays, prs, table):
    y_top, y_base, mu, std = arrays
    po.slice_forward(y_top, y_base, mu, std, prs, table, want=("mask", "y_hat", "lik", "idx"))


def torch_reference_throughput(wl, seconds_budget):
    """When PIC_REFERENCE_ROOT names a checkout of the reference (the build container; never the GPU box, where the
    variable is unset and nothing is read) the reference's own torch functions are timed too: kind = "reference"."""
    root = os.environ.get("PIC_REFERENCE_ROOT")
    if not root or not os.path.isdir(root):
        return None
    try:
        import torch
        import ref_shim

        if not ref_shim.reference_available():
            return None
        ref = ref_shim.load_reference()
        torch.set_num_threads(host_threads())
        n = wl["n"]
        units = max(1, min(8, (1 << 21) // n))
        hw = n // 32
        arrays = [torch.from_numpy(a).reshape(units, 32, hw // 16, 16) for a in make_host_inputs(n, units, 99)]
        masking, gc = ref.ChannelMask("point-based-std"), ref_shim.make_gaussian_conditional(ref)
        t0 = time.perf_counter()
        passes = 0
        while time.perf_counter() - t0 < seconds_budget and passes < 20:
            ref_shim.reference_slice_forward(ref, gc, masking, *arrays, 5.0)
            passes += 1
        dt = (time.perf_counter() - t0) / max(passes, 1)
        return {"value": units * n / dt / 1e9, "unit": "Gelem/s", "cores": torch.get_num_threads(), "kind": "reference",
                "sample": f"{units} units x {n} elem through the reference's own torch functions ({passes} passes)"}
    except Exception as exc:  # pragma: no cover - depends on the container
        return {"unavailable": f"{type(exc).__name__}: {exc}"}


def cpu_reference_throughput(wl, seconds_budget):
    """Times the oracle port (oracle/pic_oracle.c, OpenMP over units) on the units of one full step."""
    import pic_oracle as po

    po.build()
    po.set_num_threads(host_threads())
    cores = po.num_threads()
    n = wl["n"]
    step_units, prs, total = cpu_step_units(wl)
    arrays = make_host_inputs(n, step_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    cpu_pass(po, arrays, prs, table)  # warm-up
    times, t_start = [], time.perf_counter()
    while True:
        t0 = time.perf_counter()
        cpu_pass(po, arrays, prs, table)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > seconds_budget or len(times) >= 50:
            break
    med = statistics.median(times)
    out = {"value": step_units * n / med / 1e9, "unit": "Gelem/s", "cores": cores, "kind": "port",
           "sample": f"{step_units} of the step's {total} units x {n} elem (forward, codec outputs), median of {len(times)} "
                     f"passes of oracle/pic_oracle.c (qsort quantile + erfc + 63-step index), {cores} OpenMP threads"}
    tr = torch_reference_throughput(wl, 5.0)
    if tr is not None:
        out["torch_reference"] = tr
    return out


def run_reference(args, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import pic_oracle as po

    po.build()
    po.set_num_threads(host_threads())
    cores = po.num_threads()
    n = wl["n"]
    step_units, prs, total = cpu_step_units(wl)
    arrays = make_host_inputs(n, step_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_pass(po, arrays, prs, table)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(po, arrays, prs, table)
    dt = time.perf_counter() - t0
    value = args.steps * step_units * n / dt / 1e9
    slices = wl["slices"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    sample = (f"each step = {step_units} of the {total} units x {n} elem of one {args.workload} step (forward, codec outputs) "
              f"through oracle/pic_oracle.c, the C port of the reference's torch CPU path (the reference is Python "
              f"and cannot travel to this box), {cores} OpenMP threads")
    # same `config` keys as the CUDA arm (same workload, same units per step)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"], "n_per_unit": n, "slices": slices,
                                            "units_per_step_per_gpu": per_slice * slices,
                                            "units_per_step": step_units, "elements_per_step": step_units * n},
            "cpu_baseline": {"value": value, "unit": "Gelem/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ CUDA arm
class Ctx:
    """Per-process state of the CUDA arm."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import pic_b200
        from pic_b200 import distributed as pdist
        from pic_b200 import ops

        self.torch, self.dist, self.pic, self.pdist, self.ops, self.args = torch, dist, pic_b200, pdist, ops, args
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU path)")
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.L = pic_b200.lib()
        self.table = pic_b200.get_scale_table().to(self.dev)
        self.hbm, self.peak_source = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def plan(self, n, units, needs_select=1):
        k = ctypes.c_int(0)
        kind = self.L.pic_slice_forward_plan(n, units, needs_select, ctypes.byref(k))
        return kind, k.value

    def select_counters(self):
        s, f = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
        self.L.pic_debug_select_counters(ctypes.byref(s), ctypes.byref(f))
        return int(s.value), int(f.value)

    def max_over_ranks(self, x: float) -> float:
        if self.world == 1:
            return x
        t = self.torch.tensor([x], device=self.dev, dtype=self.torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def gather(self, obj):
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out


def ev_time(torch, fn, reps=20):
    """Average device time of fn over `reps` back-to-back calls (CUDA events on the current stream)."""
    fn()
    torch.cuda.synchronize()
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for _ in range(reps):
        fn()
    a1.record()
    torch.cuda.synchronize()
    return a0.elapsed_time(a1) / reps


def timed_steps(cx: Ctx, step, steps, warmup, graph=True, windows=5, sampler=None):
    """W warm-up steps, then EXACTLY `steps` steps between barrier + synchronize (CUDA events, max over ranks);
    then `windows` more windows of the same length for the spread.  The timed steps replay a CUDA graph of one step
    when `graph` (the eager time of the same K steps is reported too)."""
    torch = cx.torch
    for _ in range(max(warmup, 3)):
        step()
    cx.barrier()
    run_step, eager_ms = step, None
    if graph:
        eager_ms = cx.max_over_ranks(ev_time(torch, step, steps))
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            step()
        for _ in range(3):
            g.replay()
        run_step = g.replay
        cx.barrier()
    t_wall0 = time.perf_counter()
    if sampler is not None and cx.rank == 0:
        sampler.start()
        time.sleep(0.1)
    cx.barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(steps):
        run_step()
    ev1.record()
    cx.barrier()
    t_wall1 = time.perf_counter()
    own_ms = ev0.elapsed_time(ev1) / steps
    ms = cx.max_over_ranks(own_ms)
    clocks = sampler.stop(t_wall0, t_wall1) if (sampler is not None and cx.rank == 0) else None
    spread = []
    for _ in range(windows):
        cx.barrier()
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        for _ in range(steps):
            run_step()
        w1.record()
        cx.barrier()
        spread.append(cx.max_over_ranks(w0.elapsed_time(w1) / steps))
    return {"ms": ms, "own_ms": own_ms, "eager_ms": eager_ms, "clocks": clocks,
            "windows": {"n": windows, "steps_each": steps, "ms_per_step": [round(v, 5) for v in spread],
                        "median": statistics.median(spread) if spread else None,
                        "min": min(spread) if spread else None, "max": max(spread) if spread else None}}


def roofline(cx: Ctx, name, kernel, kern_ms, per_launch_elems, kbytes, elems_per_rank, step_bytes_per_elem, ms_per_step,
             other=None):
    bytes_per_launch = per_launch_elems * kbytes
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(name)
        except Exception:
            traffic = None
    return {"bound": "hbm", "achieved": achieved, "peak": cx.hbm, "unit": "GB/s", "frac": achieved / cx.hbm,
            "traffic": traffic, "peak_source": cx.peak_source, "kernel": kernel,
            "algorithmic_bytes_per_elem": kbytes, "algorithmic_bytes_per_launch": bytes_per_launch,
            "launch_ms": kern_ms, "other_kernels": other,
            "whole_step_frac": (elems_per_rank * step_bytes_per_elem / (ms_per_step * 1e-3) / 1e9) / cx.hbm}


# ---- kodak_sweep (headline) + per_slice + sweep_shared -------------------------------------------------------------
def bench_kodak(cx: Ctx, sampler):
    torch, ops, args = cx.torch, cx.ops, cx.args
    wl = WORKLOADS["kodak_sweep"]
    n, slices = wl["n"], wl["slices"]
    per_slice = len(wl["prs"])
    units = per_slice * slices
    seed = 1234 + 1000 * 2 + cx.rank
    q_slice = ops.q01_tensor(unit_prs(wl, per_slice, 4321 + cx.rank), cx.dev)
    q_all = torch.cat([q_slice] * slices).contiguous()
    y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, cx.dev)
    want = ("mask", "y_hat", "lik", "idx")
    outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=cx.dev) for k in want}
    ws = torch.empty(ops.workspace_bytes(n, units), dtype=torch.uint8, device=cx.dev)
    step_kind, step_k = cx.plan(n, units)

    def step():
        ops.slice_forward(y_top, y_base, mu, std, units, q_all, cx.table, want=want, out=outs, workspace=ws)

    c0 = cx.select_counters()
    t = timed_steps(cx, step, args.steps, args.warmup, graph=bool(args.graph), sampler=sampler)
    c1 = cx.select_counters()
    ms = t["ms"]
    elems_per_rank, total_elems = units * n, units * n * cx.world
    value = total_elems / (ms * 1e-3) / 1e9
    # dominant kernel (apply, thresholds given) and the select kernel alone, CUDA events on this stream
    thr0 = ops.select_threshold(std, units, q_all, workspace=ws)
    apply_ms = ev_time(torch, lambda: ops.slice_forward(y_top, y_base, mu, std, units, q_all, cx.table, thr_in=thr0,
                                                        want=want, out=outs, workspace=ws))
    sel_ms = ev_time(torch, lambda: ops.select_threshold(std, units, q_all, workspace=ws))
    other = {"select_kernel": "select_lean_kernel", "select_ms": sel_ms,
             "select_GBps_of_std": units * n * 4 / (sel_ms * 1e-3) / 1e9, "select_frac_of_hbm": units * n * 4 / (sel_ms * 1e-3) / 1e9 / cx.hbm,
             "apply_share_of_step": apply_ms / t["own_ms"]}
    roof = roofline(cx, "kodak_sweep", "slice_apply_kernel", apply_ms, units * n, 32, elems_per_rank, 32, t["own_ms"], other)
    per_rank = cx.gather({"rank": cx.rank, "seed": seed, "ms_per_step": round(t["own_ms"], 5), "select_ms": round(sel_ms, 5),
                          "apply_ms": round(apply_ms, 5), "select_units_sampled": c1[0] - c0[0],
                          "select_units_fallback": c1[1] - c0[1]})

    # ---- end to end through host buffers (before the side workloads overwrite `outs`)
    e2e = None
    if not args.no_e2e:
        e2e = bench_e2e(cx, n, units, per_slice, (y_top, y_base, mu, std), q_all, outs, total_elems)

    # ---- encoder order: one select + one apply launch per slice index
    def view(tn, s):
        return tn[s * per_slice:(s + 1) * per_slice]

    ws_s = torch.empty(ops.workspace_bytes(n, per_slice), dtype=torch.uint8, device=cx.dev)

    def step_per_slice():
        for s in range(slices):
            ops.slice_forward(view(y_top, s), view(y_base, s), view(mu, s), view(std, s), per_slice, q_slice, cx.table,
                              want=want, out={k: view(v, s) for k, v in outs.items()}, workspace=ws_s)

    extra = {}
    if not args.only_main:
        tp = timed_steps(cx, step_per_slice, max(5, args.steps // 2), 3, graph=bool(args.graph), windows=3)
        _, ps_k = cx.plan(n, per_slice)
        extra["per_slice"] = {
            "desc": "kodak_sweep units in encoder order: 10 dependent slices, one select + one apply launch per slice (101 units each)",
            "value": total_elems / (tp["ms"] * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": tp["ms"], "windows": tp["windows"],
            "launches_per_step": slices * ps_k, "bytes_per_elem": 32,
            "whole_step_frac": (elems_per_rank * 32 / (tp["own_ms"] * 1e-3) / 1e9) / cx.hbm, "scaling": "weak"}
        # ---- the sweep on ONE set of latents per slice (what check_levels_np evaluates): inputs [slices, n]
        yt1, yb1, mu1, sd1 = (view(a, 0)[:slices].contiguous() for a in (y_top, y_base, mu, std))
        q_lv = torch.cat([q_slice] * 1).contiguous()
        outs_m = {k: v.view(slices, per_slice, n) for k, v in outs.items()}
        outs_m["thr"] = torch.empty((slices, per_slice), dtype=torch.float32, device=cx.dev)

        def step_shared():
            for s in range(slices):   # slices stay sequential (each slice's mu / std depend on the previous ones)
                ops.slice_forward_multi(yt1[s:s + 1], yb1[s:s + 1], mu1[s:s + 1], sd1[s:s + 1], 1, wl["prs"], cx.table,
                                        want=want, out={k: v[s:s + 1] for k, v in outs_m.items()}, q01_levels=q_lv)

        tsd = timed_steps(cx, step_shared, max(5, args.steps // 2), 3, graph=bool(args.graph), windows=3)
        extra["sweep_shared"] = {
            "desc": "quality sweep of ONE set of Kodak latents per slice: 10 sequential slices x 101 qualities sharing the "
                    "slice's inputs (pic_slice_forward_multi), codec outputs for every (slice, quality)",
            "value": total_elems / (tsd["ms"] * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": tsd["ms"], "windows": tsd["windows"],
            "launches_per_step": slices * 2, "bytes_per_elem": 16,
            "whole_step_frac": (elems_per_rank * 16 / (tsd["own_ms"] * 1e-3) / 1e9) / cx.hbm, "scaling": "weak",
            "note": "algorithmic bytes: 16 B/element of output; the shared inputs cross HBM once per slice"}

    cfg = {"workload": "kodak_sweep", "desc": wl["desc"], "n_per_unit": n, "slices": slices,
           "units_per_step_per_gpu": units, "elements_per_step": total_elems, "launch": "per_step",
           "cuda_graph": bool(args.graph), "launches_per_step": step_k, "outputs": list(want), "bytes_per_elem": 32,
           "l2": f"inputs of one step = {elems_per_rank * 16 / 1e6:.0f} MB per GPU > 126 MB L2, no explicit flush",
           "parallelism": f"{cx.world} rank(s), units sharded, no collective"}
    del y_top, y_base, mu, std, outs
    torch.cuda.empty_cache()
    return {"value": value, "ms": ms, "t": t, "roofline": roof, "per_rank": per_rank, "config": cfg, "extra": extra,
            "e2e": e2e, "launches": step_k}


def bench_e2e(cx: Ctx, n, units, per_slice, inputs, q_all, outs, total_elems):
    """The same step through pic_slice_forward_host: pinned host buffers, H2D + kernels + D2H inside the timed region.
    Two output formats: the API's f32 / i32 arrays, and compact u8 mask + u8 index (10 B/element back instead of 16)."""
    torch, ops, args, L = cx.torch, cx.ops, cx.args, cx.L
    chunk = args.e2e_chunk if args.e2e_chunk > 0 else max(1, min(per_slice, (48 << 20) // (n * 4)))
    all_cpus = os.sched_getaffinity(0)
    numa_node = None if args.no_numa_bind else ops.bind_host_to_device_numa(cx.dev)
    host_in = [t.cpu().pin_memory() for t in inputs]
    q_host = q_all.cpu().pin_memory()
    want = ("mask", "y_hat", "lik", "idx")
    host_out = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32).pin_memory() for k in want}
    nbytes = int(L.pic_host_pipeline_bytes(n, chunk))
    dbuf = torch.empty(nbytes, dtype=torch.uint8, device=cx.dev)
    tb = cx.table.cpu()
    torch.cuda.synchronize()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def run(fn):
        fn()
        cx.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()          # returns when the outputs are in host memory
        cx.barrier()
        return cx.max_over_ranks(time.perf_counter() - t0)

    def host_step():
        rc = L.pic_slice_forward_host(host_in[0].data_ptr(), host_in[1].data_ptr(), host_in[2].data_ptr(),
                                      host_in[3].data_ptr(), 0.5, q_host.data_ptr(), None, tb.data_ptr(), 64,
                                      0.11, 1e-9, n, units, chunk, host_out["mask"].data_ptr(),
                                      host_out["y_hat"].data_ptr(), host_out["lik"].data_ptr(),
                                      host_out["idx"].data_ptr(), None, None, None, dbuf.data_ptr(), nbytes)
        if rc != 0:
            raise RuntimeError(f"pic_slice_forward_host rc={rc}")

    dt = run(host_step)
    if cx.rank == 0:
        assert torch.equal(host_out["mask"][:per_slice], outs["mask"][:per_slice].cpu()), "host/device mismatch"
    # plain pinned copies of the same byte counts, both directions at once: the PCIe ceiling of this box
    h2d_b, d2h_b = units * n * 16, units * n * 16
    dsrc = torch.empty(h2d_b // 4, dtype=torch.float32, device=cx.dev)
    hsrc = torch.empty(h2d_b // 4, dtype=torch.float32).pin_memory()
    hdst = torch.empty(d2h_b // 4, dtype=torch.float32).pin_memory()
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def copies():
        with torch.cuda.stream(s_in):
            dsrc.copy_(hsrc, non_blocking=True)
        with torch.cuda.stream(s_out):
            hdst.copy_(dsrc, non_blocking=True)
        s_in.synchronize()
        s_out.synchronize()

    dt_copy = run(copies)
    del dsrc, hsrc, hdst
    e2e = {"value": total_elems * e2e_steps / dt / 1e9, "unit": "Gelem/s",
           "h2d_bytes_per_step": int(units * n * 16 + units * 4), "d2h_bytes_per_step": int(units * n * 16),
           "steps": e2e_steps, "value_scope": "whole job (all ranks' elements / slowest rank's wall time)",
           "bytes_scope": "per rank and step", "host_numa_node": numa_node,
           "api": "pic_slice_forward_host (C ABI, pinned host buffers, 3-slot copy/compute pipeline)",
           "pcie_copy_only": {"value": total_elems * e2e_steps / dt_copy / 1e9, "unit": "Gelem/s",
                              "what": "cudaMemcpyAsync H2D || D2H of the same byte counts, no kernels (pinned, two streams)"},
           "frac_of_copy_only": dt_copy / dt}
    # compact outputs: u8 mask + u8 index
    if hasattr(L, "pic_slice_forward_host_compact"):
        hm = torch.empty((units, n), dtype=torch.uint8).pin_memory()
        hi = torch.empty((units, n), dtype=torch.uint8).pin_memory()
        nbytes_c = int(L.pic_host_pipeline_bytes(n, chunk))

        def host_step_c():
            rc = L.pic_slice_forward_host_compact(host_in[0].data_ptr(), host_in[1].data_ptr(), host_in[2].data_ptr(),
                                                  host_in[3].data_ptr(), 0.5, q_host.data_ptr(), tb.data_ptr(), 64, 0.11, 1e-9,
                                                  n, units, chunk, hm.data_ptr(), host_out["y_hat"].data_ptr(),
                                                  host_out["lik"].data_ptr(), hi.data_ptr(), dbuf.data_ptr(), nbytes_c)
            if rc != 0:
                raise RuntimeError(f"pic_slice_forward_host_compact rc={rc}")

        dtc = run(host_step_c)
        if cx.rank == 0:
            assert torch.equal(hm[:per_slice].float(), outs["mask"][:per_slice].cpu()), "compact mask mismatch"
            assert torch.equal(hi[:per_slice].int(), outs["idx"][:per_slice].cpu()), "compact index mismatch"
        e2e["compact"] = {"value": total_elems * e2e_steps / dtc / 1e9, "unit": "Gelem/s",
                          "h2d_bytes_per_step": int(units * n * 16 + units * 4), "d2h_bytes_per_step": int(units * n * 10),
                          "api": "pic_slice_forward_host_compact (u8 mask, u8 scale index, f32 y_hat / likelihood)"}
    os.sched_setaffinity(0, all_cpus)   # the CPU baseline gets every host core again
    return e2e


# ---- first_train ---------------------------------------------------------------------------------------------------
def bench_first_train(cx: Ctx):
    torch, ops, args = cx.torch, cx.ops, cx.args
    wl = WORKLOADS["first_train"]
    n, slices, per_slice = wl["n"], wl["slices"], wl["batch"]
    units = per_slice * slices
    seed = 1234 + 1000 * 3 + cx.rank
    q_all = torch.cat([ops.q01_tensor(unit_prs(wl, per_slice, 4321 + cx.rank), cx.dev)] * slices).contiguous()
    y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, cx.dev)
    g = torch.Generator(device=cx.dev).manual_seed(seed + 7)
    noise = torch.rand((units, n), device=cx.dev, generator=g) - 0.5
    g_lik = torch.randn((units, n), device=cx.dev, generator=g)
    g_yhat = torch.randn((units, n), device=cx.dev, generator=g)
    want = ("mask", "y_hat", "lik")
    outs = {k: torch.empty((units, n), dtype=torch.float32, device=cx.dev) for k in want}
    ws = torch.empty(ops.workspace_bytes(n, units), dtype=torch.uint8, device=cx.dev)
    fwd_kind, fwd_k = cx.plan(n, units)

    def fwd():
        ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, noise=noise, want=want, out=outs, workspace=ws)

    def step():
        fwd()
        ops.slice_backward(g_lik, g_yhat, y_top, y_base, mu, std, outs["mask"], noise)

    t = timed_steps(cx, step, max(5, args.steps // 2), 3, graph=False, windows=3)
    fwd_ms = ev_time(torch, fwd, 10)
    bwd_ms = ev_time(torch, lambda: ops.slice_backward(g_lik, g_yhat, y_top, y_base, mu, std, outs["mask"], noise), 10)
    elems = units * n
    rec = {"desc": wl["desc"], "value": elems * cx.world / (t["ms"] * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": t["ms"],
           "windows": t["windows"], "launches_per_step": fwd_k + 1, "bytes_per_elem": wl["bytes_per_elem"], "scaling": "weak",
           "whole_step_frac": (elems * wl["bytes_per_elem"] / (t["own_ms"] * 1e-3) / 1e9) / cx.hbm,
           "forward_kernel": {"name": "slice_fused_kernel (training forward)", "ms": fwd_ms,
                              "frac": (elems * 32 / (fwd_ms * 1e-3) / 1e9) / cx.hbm, "bytes_per_elem": 32},
           "backward_kernel": {"name": "slice_backward_kernel", "ms": bwd_ms,
                               "frac": (elems * 48 / (bwd_ms * 1e-3) / 1e9) / cx.hbm, "bytes_per_elem": 48,
                               "note": "timed back to back on the same tensors: each of the 8 input arrays is 84 MB, so part of "
                                       "them is still L2-resident from the previous launch and the fraction of the HBM peak can "
                                       "exceed what DRAM alone delivers; inside the step it is ms_per_step - forward ms = "
                                       f"{t['own_ms'] - fwd_ms:.4f} ms"}}
    del y_top, y_base, mu, std, noise, g_lik, g_yhat, outs
    torch.cuda.empty_cache()
    return rec


# ---- rem_latent ----------------------------------------------------------------------------------------------------
def bench_rem(cx: Ctx):
    torch, ops, args = cx.torch, cx.ops, cx.args
    wl = WORKLOADS["rem_latent"]
    n, slices, per_slice = wl["n"], wl["slices"], wl["batch"]
    units = per_slice * slices
    seed = 1234 + 1000 * 4 + cx.rank
    q_all = torch.cat([ops.q01_tensor(unit_prs(wl, per_slice, 4321 + cx.rank), cx.dev)] * slices).contiguous()
    # models/rem_pic.py:181-195 (star mask on the pre-REM scale, duplicated for mu_std), 121-132 (checkpoint
    # representation at quality_ref = 0.75 -> a second select), 382-391 (block mask on the refined scale + slice)
    y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, cx.dev)
    _, _, _, std_pre = make_device_inputs(torch, n, units, seed + 17, cx.dev)
    want = ("mask", "y_hat", "lik")
    outs = {k: torch.empty((units, n), dtype=torch.float32, device=cx.dev) for k in want}
    att = torch.empty((units, 2, n), dtype=torch.float32, device=cx.dev)
    ws = torch.empty(ops.workspace_bytes(n, units), dtype=torch.uint8, device=cx.dev)
    _, fwd_k = cx.plan(n, units)
    q_ckpt = ops.pr_to_q01(0.75)

    def fwd():
        ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, want=want, out=outs, workspace=ws)

    def step():
        ops.attention_mask(std_pre, units, q_all, copies=2, out=att, workspace=ws)   # pre-REM star mask, cat([m, m], 1)
        ops.select_threshold(std_pre, units, q_ckpt, workspace=ws)                  # checkpoint pass threshold
        fwd()

    t = timed_steps(cx, step, max(5, args.steps // 2), 3, graph=False, windows=3)
    fwd_ms = ev_time(torch, fwd, 10)
    elems = units * n
    rec = {"desc": wl["desc"], "value": elems * cx.world / (t["ms"] * 1e-3) / 1e9, "unit": "Gelem/s", "ms_per_step": t["ms"],
           "windows": t["windows"], "launches_per_step": 3 + fwd_k, "bytes_per_elem": wl["bytes_per_elem"], "scaling": "weak",
           "whole_step_frac": (elems * wl["bytes_per_elem"] / (t["own_ms"] * 1e-3) / 1e9) / cx.hbm,
           "forward_kernel": {"name": "slice_fused_kernel (codec forward, 3 outputs)", "ms": fwd_ms,
                              "frac": (elems * 28 / (fwd_ms * 1e-3) / 1e9) / cx.hbm, "bytes_per_elem": 28}}
    del y_top, y_base, mu, std, std_pre, outs, att
    torch.cuda.empty_cache()
    return rec


# ---- tile8192 ------------------------------------------------------------------------------------------------------
def bench_tile(cx: Ctx):
    """One 8192x8192 image: N = 1 runs the large-unit select on one GPU; N > 1 splits every slice into row bands and
    the thresholds come from the library-issued NCCL histogram all-reduces (a real collective on the data path)."""
    torch, ops, pdist, args = cx.torch, cx.ops, cx.pdist, cx.args
    wl = WORKLOADS["tile8192"]
    n, units, world = wl["n"], wl["slices"], cx.world
    n_local = n // world
    seed = 1234 + 1000 * 5 + cx.rank
    q_all = ops.q01_tensor([1.0] * units, cx.dev)
    y_top, y_base, mu, std = make_device_inputs(torch, n_local, units, seed, cx.dev)
    want = ("mask", "y_hat", "lik", "idx")
    outs = {k: torch.empty((units, n_local), dtype=torch.int32 if k == "idx" else torch.float32, device=cx.dev) for k in want}
    rec = {"desc": wl["desc"], "unit": "Gelem/s", "bytes_per_elem": wl["bytes_per_elem"], "scaling": "strong",
           "n_per_unit": n, "n_local_per_rank": n_local}
    comm = None
    if world == 1:
        ws = torch.empty(ops.workspace_bytes(n, units), dtype=torch.uint8, device=cx.dev)

        def step():
            ops.slice_forward(y_top, y_base, mu, std, units, q_all, cx.table, want=want, out=outs, workspace=ws)
        rec["launches_per_step"] = cx.plan(n, units)[1]
        rec["collectives_per_step"] = 0
        graph = bool(args.graph)
    else:
        backend = pdist.CudaTileBackend(std, units)
        comm = pdist.NcclTileComm(cx.dev)
        t_c = pdist.tiled_select_threshold(std, units, n, q_all, comm=comm)
        t_t = pdist.tiled_select_threshold(std, units, n, q_all, backend=backend)
        same = bool(torch.equal(t_c, t_t))
        every = [torch.zeros_like(t_c) for _ in range(world)]
        cx.dist.all_gather(every, t_c)
        same_ranks = all(bool(torch.equal(every[0], e)) for e in every)
        rec["thresholds_equal"] = {"library_nccl_vs_torch_distributed": same, "across_ranks": same_ranks}

        protocol = os.environ.get("PIC_TILED_PROTOCOL", "p2p")
        t_s = pdist.tiled_select_threshold(std, units, n, q_all, comm=comm, protocol="sampled")
        rec["thresholds_equal"]["sampled_vs_rounds"] = bool(torch.equal(t_s, t_c))
        if protocol == "p2p":
            try:
                comm.enable_p2p(n, units)
                t_p = comm.select_threshold(std, units, n, q_all, protocol="p2p")
                rec["thresholds_equal"]["p2p_vs_rounds"] = bool(torch.equal(t_p, t_c))
            except Exception as exc:  # no CUDA IPC between the ranks (all ranks fail together): NCCL carries the exchange
                rec["p2p_unavailable"] = f"{type(exc).__name__}: {exc}"
                protocol = "rounds"

        def step():
            if protocol == "p2p":
                thr = comm.select_threshold(std, units, n, q_all, protocol="p2p", check_status=False)
            else:
                thr = pdist.tiled_select_threshold(std, units, n, q_all, comm=comm, protocol=protocol)
            ops.slice_forward(y_top, y_base, mu, std, units, q_all, cx.table, thr_in=thr, want=want, out=outs)
        rec["protocol"] = protocol
        if protocol == "p2p":
            rec["launches_per_step"] = 8     # sample+exchange, pivot select, pivots, sweep, exchange(+pack), merge, cluster select, apply
            rec["collectives_per_step"] = 0
            rec["collective"] = ("two peer-memory exchanges (NVLink stores into the peers' CUDA IPC windows + flags) issued by "
                                 "libpic_latent.so's own kernels: the bands' samples, then bracket counts + candidates; no NCCL "
                                 "call, no host synchronisation, validity word checked after the timed region")
            graph = bool(args.graph) and os.environ.get("PIC_TILED_GRAPH", "1") != "0"
            comm._p2p_status.zero_()
        elif protocol == "sampled":
            rec["launches_per_step"] = 9     # sample, pool, pivot select, pivots, sweep, pack, merge, cluster select, apply
            rec["collectives_per_step"] = 2
            rec["collective"] = ("ncclAllGather of the bands' samples + ncclAllGather of the bracket counts and candidates, issued "
                                 "by libpic_latent.so (one stream synchronisation per select; histogram rounds only on a miss)")
            graph = False
        else:
            rec["launches_per_step"] = 9
            rec["collectives_per_step"] = 4
            rec["collective"] = "ncclAllReduce(uint32 sum) of the per-slice radix histograms + ncclAllReduce(min), issued by libpic_latent.so"
            # the collectives belong to the library's own communicator, so the whole step (kernels + NCCL) can be captured
            # in a CUDA graph; PIC_TILED_GRAPH=0 keeps it eager
            graph = bool(args.graph) and os.environ.get("PIC_TILED_GRAPH", "1") != "0"
        rec["cuda_graph"] = graph
    try:
        t = timed_steps(cx, step, max(5, args.steps // 2), 3, graph=graph, windows=3)
    except Exception:
        if comm is not None:
            comm.close()       # collective teardown: every rank raises together (same step, same inputs shape)
        raise
    rec.update({"value": units * n / (t["ms"] * 1e-3) / 1e9, "ms_per_step": t["ms"], "windows": t["windows"],
                "ms_per_step_by_rank": [round(v, 5) for v in cx.gather(t["own_ms"])],
                "whole_step_frac": (units * n_local * wl["bytes_per_elem"] / (t["own_ms"] * 1e-3) / 1e9) / cx.hbm})
    if comm is not None:
        rec["sampled_selects_that_fell_back"] = comm.fallbacks
        if rec.get("protocol") == "p2p":
            rec["p2p_status_after_timed_region"] = comm.p2p_status()     # 0: every bracket held, no exchange timed out
        comm.close()
    del y_top, y_base, mu, std, outs
    torch.cuda.empty_cache()
    return rec


def run_cuda(args):
    # stdout carries exactly ONE JSON line: anything libraries print to fd 1 (e.g. the NCCL version banner) is sent
    # to stderr, the JSON line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    cx = Ctx(args)
    sampler = ClockSampler(cx.local_rank, cx.world)
    main = bench_kodak(cx, sampler)
    workloads = dict(main["extra"])
    if not args.only_main:
        for name, fn in (("first_train", bench_first_train), ("rem_latent", bench_rem), ("tile8192", bench_tile)):
            try:
                workloads[name] = fn(cx)
            except Exception as exc:  # a failing side workload must not take the headline down
                workloads[name] = {"error": f"{type(exc).__name__}: {exc}"}
                cx.torch.cuda.empty_cache()
    if cx.rank != 0:
        if cx.world > 1:
            cx.dist.destroy_process_group()
        return
    wl = WORKLOADS["kodak_sweep"]
    cpu = None if args.no_cpu else cpu_reference_throughput(wl, args.cpu_seconds)
    t = main["t"]
    line = {
        "metric": METRIC, "value": main["value"], "unit": "Gelem/s", "n_gpus": cx.world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": main["ms"], "higher_is_better": True, "scaling": wl["scaling"],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "ms_per_step_by_rank": [r["ms_per_step"] for r in main["per_rank"]], "per_rank": main["per_rank"],
        "windows": t["windows"], "eager_ms_per_step": t["eager_ms"], "config": main["config"],
        "roofline": main["roofline"], "cpu_baseline": cpu, "e2e": main["e2e"],
        "gpu_launches": main["launches"] * args.steps, "clocks": t["clocks"], "workloads": workloads,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if cx.world > 1:
        cx.dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="kodak_sweep", choices=sorted(WORKLOADS),
                    help="reference arm: which configuration the CPU path runs (the CUDA arm always prints kodak_sweep "
                         "with the other configurations in its `workloads` block)")
    ap.add_argument("--only-main", action="store_true", help="skip the `workloads` block")
    ap.add_argument("--graph", type=int, default=1, help="1: timed steps replay a CUDA graph of one step (default; NCCL steps stay eager)")
    ap.add_argument("--no-numa-bind", action="store_true", help="e2e: do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="units per pipeline chunk of the host-buffer path (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, WORKLOADS[args.workload])
    else:
        run_cuda(args)


if __name__ == "__main__":
    main()
