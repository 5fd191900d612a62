#!/usr/bin/env python
"""bench.py -- mask+quantize+likelihood throughput of the PIC latent hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N>1: one rank per GPU

Workloads (--workload):
  kodak_sweep (default, BASELINE config[1]): Kodak-shape 768x512 latents ([1,32,32,48] => n=49152 per
      unit), 10 progressive slices x 101-point quality sweep (pr = 0, 0.1, .. 10) = 1010 units per step,
      codec outputs (mask, y_hat, likelihood, scale index).  One step = ONE launch of the fused kernel over
      all 1010 (slice, quality) units (--launch per_slice: one launch per slice index).  Weak scaling:
      every rank runs a full replica on its own data, no collective on the data path.
  first_train (config[2]): [256,32,16,16] x 10 slices, random quality per image, training-mode forward
      (noise) + fused backward.  Weak scaling, no collective.
  rem_latent (config[3]): REM variant of the path at the first_train shape: 3 threshold selections per slice
      (pre-REM attention mask duplicated to [B,64,h,w], checkpoint pass at q = 0.75, post-REM block mask) + slice.
  tile8192 (config[4]): one 8192x8192 image, 10 slices of n=8388608, each rank holds a row band; the
      per-slice threshold comes from NCCL all-reduced radix histograms.  Strong scaling.
Inputs of one step are far larger than the 126 MB L2 (kodak_sweep: 794 MB), so consecutive steps stream
from HBM without an explicit flush.  One JSON line is printed by rank 0 (keys: DESIGN.md "Measurement").
"""
from __future__ import annotations

import argparse
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
                     desc="single 8192x8192 image, 10 slices, q=1, row bands over the ranks, NCCL histogram all-reduce",
                     bytes_per_elem=36,  # std read by the select rounds (>=1 pass from HBM) + the 32 of the apply
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
        use = inside if inside else anytime
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


def cpu_sample(wl, cores):
    """Bounded sample of the workload for the CPU arm: same unit size, a few units per host thread."""
    n = wl["n"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    total = per_slice * wl["slices"]
    sample_units = max(1, min(total, max(cores * 2, (4 << 20) // n)))
    if n > (1 << 20):
        sample_units = 1          # one 8M-element unit is ~1 s of qsort per pass
    prs_all = unit_prs(wl, per_slice, 4321)
    prs = [prs_all[u % len(prs_all)] for u in range(sample_units)]
    return sample_units, prs


def cpu_pass(po, wl, arrays, prs, table):
    y_top, y_base, mu, std = arrays
    po.slice_forward(y_top, y_base, mu, std, prs, table, want=("mask", "y_hat", "lik", "idx"))


def cpu_reference_throughput(wl, seconds_budget):
    """Times the oracle port (oracle/pic_oracle.c, OpenMP over units) on a bounded sample."""
    import pic_oracle as po

    po.build()
    po.set_num_threads(host_threads())
    cores = po.num_threads()
    n = wl["n"]
    sample_units, prs = cpu_sample(wl, cores)
    arrays = make_host_inputs(n, sample_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    cpu_pass(po, wl, arrays, prs, table)  # warm-up
    times, t_start = [], time.perf_counter()
    while True:
        t0 = time.perf_counter()
        cpu_pass(po, wl, arrays, prs, table)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > seconds_budget or len(times) >= 50:
            break
    med = statistics.median(times)
    return {"value": sample_units * n / med / 1e9, "unit": "Gelem/s", "cores": cores, "kind": "port",
            "sample": f"{sample_units} units x {n} elem (forward, codec outputs), median of {len(times)} passes of "
                      f"oracle/pic_oracle.c (qsort quantile + erfc + 63-step index), {cores} OpenMP threads"}


def run_reference(args, wl):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    import pic_oracle as po

    po.build()
    po.set_num_threads(host_threads())
    cores = po.num_threads()
    n = wl["n"]
    sample_units, prs = cpu_sample(wl, cores)
    arrays = make_host_inputs(n, sample_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    for _ in range(args.warmup):
        cpu_pass(po, wl, arrays, prs, table)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_pass(po, wl, arrays, prs, table)
    dt = time.perf_counter() - t0
    value = args.steps * sample_units * n / dt / 1e9
    sample = (f"each step = {sample_units} units x {n} elem of the {args.workload} workload (forward, codec outputs) "
              f"through oracle/pic_oracle.c, the C port of the reference's torch CPU path (the reference is Python "
              f"and cannot travel to this box), {cores} OpenMP threads")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": {"workload": args.workload, "desc": wl["desc"], "n_per_unit": n,
                                            "units_per_step": sample_units},
            "cpu_baseline": {"value": value, "unit": "Gelem/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Gelem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ CUDA arm
def run_cuda(args, wl):
    import torch
    import torch.distributed as dist

    import pic_b200
    from pic_b200 import distributed as pdist
    from pic_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # stdout carries exactly ONE JSON line: anything libraries print to fd 1 (e.g. the NCCL version
    # banner) is sent to stderr, the JSON line is written to the saved descriptor at the end.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = pic_b200.lib()
    name = args.workload
    n, slices = wl["n"], wl["slices"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    units = per_slice * slices
    prs = unit_prs(wl, per_slice, 4321 + rank)
    seed = 1234 + 1000 * 2 + rank
    table = pic_b200.get_scale_table().to(dev)
    q_slice = ops.q01_tensor(prs, dev)
    q_all = torch.cat([q_slice] * slices).contiguous()
    import ctypes

    def plan(n_, units_, needs_select=1):
        k = ctypes.c_int(0)
        kind = L.pic_slice_forward_plan(n_, units_, needs_select, ctypes.byref(k))
        return kind, k.value

    fused = n <= int(L.pic_fused_max_elems())
    launches = [0]
    tile_comm = None
    PLAN_KERNEL = {0: "slice_fused_kernel", 1: "slice_apply_kernel", 2: "slice_apply_kernel"}
    main_kernel = "slice_fused_kernel"

    if name == "tile8192":
        # strong scaling: the unit is split into `world` row bands; every rank owns one band of each slice
        n_local = n // world
        y_top, y_base, mu, std = make_device_inputs(torch, n_local, units, seed, dev)
        want = ("mask", "y_hat", "lik", "idx")
        outs = {k: torch.empty((units, n_local), dtype=torch.int32 if k == "idx" else torch.float32, device=dev)
                for k in want}
        backend = pdist.CudaTileBackend(std, units) if world > 1 else None
        if world > 1 and not args.torch_collectives:
            # collectives issued inside libpic_latent.so (one host call per select); checked once against the
            # torch.distributed protocol: thresholds must be bit-identical
            tile_comm = pdist.NcclTileComm(dev)
            t_c = pdist.tiled_select_threshold(std, units, n, q_all, comm=tile_comm)
            t_t = pdist.tiled_select_threshold(std, units, n, q_all, backend=backend)
            if not torch.equal(t_c, t_t):
                raise RuntimeError("tiled select: library-issued NCCL path disagrees with the torch.distributed path")
        main_kernel = "slice_apply_kernel" if (world == 1 or n_local > int(L.pic_fused_max_elems())) else "slice_fused_kernel"

        def step():
            if world == 1:
                ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, want=want, out=outs)
                launches[0] = plan(n, units)[1]   # pivot + sweep + cluster select + apply = 4
            else:
                thr = pdist.tiled_select_threshold(std, units, n, q_all, backend=backend, comm=tile_comm)
                ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, thr_in=thr, want=want, out=outs)
                launches[0] = 9          # begin + 3 x (hist, advance) + finish + apply; + 4 NCCL all-reduces (not counted)
        elems_per_rank = units * n_local
        total_elems = units * n
    elif name == "first_train":
        y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, dev)
        g = torch.Generator(device=dev).manual_seed(seed + 7)
        noise = torch.rand((units, n), device=dev, generator=g) - 0.5
        g_lik = torch.randn((units, n), device=dev, generator=g)
        g_yhat = torch.randn((units, n), device=dev, generator=g)
        want = ("mask", "y_hat", "lik")
        outs = {k: torch.empty((units, n), dtype=torch.float32, device=dev) for k in want}

        fwd_kind, fwd_k = plan(n, units)
        main_kernel = PLAN_KERNEL[fwd_kind]

        def step():
            ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, noise=noise, want=want, out=outs)
            ops.slice_backward(g_lik, g_yhat, y_top, y_base, mu, std, outs["mask"], noise)
            launches[0] = fwd_k + 1
        elems_per_rank = units * n
        total_elems = elems_per_rank * world
    elif name == "rem_latent":
        # models/rem_pic.py:181-195 (star mask on the pre-REM scale, duplicated for mu_std), 121-132 (checkpoint
        # representation at quality_ref = 0.75 -> a second select), 382-391 (block mask on the refined scale + slice)
        y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, dev)
        _, _, _, std_pre = make_device_inputs(torch, n, units, seed + 17, dev)
        want = ("mask", "y_hat", "lik")
        outs = {k: torch.empty((units, n), dtype=torch.float32, device=dev) for k in want}
        att = torch.empty((units, 2, n), dtype=torch.float32, device=dev)
        fwd_kind, fwd_k = plan(n, units)
        main_kernel = PLAN_KERNEL[fwd_kind]

        def step():
            ops.attention_mask(std_pre, units, q_all, copies=2, out=att)  # pre-REM star mask, cat([m, m], 1)
            ops.select_threshold(std_pre, units, ops.pr_to_q01(0.75))    # checkpoint pass threshold
            ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, want=want, out=outs)
            launches[0] = 3 + fwd_k                                      # select + mask pass, select, slice
        elems_per_rank = units * n
        total_elems = elems_per_rank * world
    else:
        y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, dev)
        want = ("mask", "y_hat", "lik", "idx")
        outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev)
                for k in want}

        def view(t, s):
            return t[s * per_slice:(s + 1) * per_slice]

        step_kind, step_k = plan(n, units) if args.launch == "per_step" else plan(n, per_slice)
        main_kernel = PLAN_KERNEL[step_kind]

        def step():
            if args.launch == "per_step":
                ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, want=want, out=outs)
                launches[0] = step_k
            else:
                for s in range(slices):
                    o = {k: view(v, s) for k, v in outs.items()}
                    ops.slice_forward(view(y_top, s), view(y_base, s), view(mu, s), view(std, s), per_slice, q_slice,
                                      table, want=want, out=o)
                launches[0] = slices * step_k
        elems_per_rank = units * n
        total_elems = elems_per_rank * world

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step()
    barrier()
    # The timed steps replay ONE step captured in a CUDA graph (the launch-bound loop: 2 launches / 0.3 ms), so that
    # host scheduling of N processes on one box does not leak into a device measurement; the same K steps issued
    # eagerly are timed first and reported as eager_ms_per_step.  NCCL steps (tiled, N > 1) stay eager.
    run_step, eager_ms = step, None
    # (capturing the NCCL step was tried at N = 2: 2.5 % faster, but the process group then hangs at teardown)
    use_graph = bool(args.graph) and not (name == "tile8192" and world > 1)
    if use_graph:
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            step()
        g1.record()
        barrier()
        eager_ms = g0.elapsed_time(g1) / args.steps
        if world > 1:
            t = torch.tensor([eager_ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            eager_ms = float(t.item())
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            step()
        for _ in range(3):
            graph.replay()
        run_step = graph.replay
        barrier()
    sampler = ClockSampler(local_rank, world)
    if rank == 0:
        sampler.start()
        time.sleep(0.1)
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        run_step()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    rank_ms = [ms / args.steps]
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        every = [torch.zeros_like(tms) for _ in range(world)]
        dist.all_gather(every, tms)
        rank_ms = [float(t.item()) / args.steps for t in every]     # reported for transparency
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)                  # the step time is the slowest rank's
        ms = float(tms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms / args.steps
    value = total_elems / (ms_per_step * 1e-3) / 1e9

    # ---------------- end-to-end: host buffers through the C ABI (H2D + kernels + D2H inside the timed region) ---
    e2e = None
    if not args.no_e2e and name == "kodak_sweep":
        chunk = args.e2e_chunk if args.e2e_chunk > 0 else max(1, min(per_slice, (48 << 20) // (n * 4)))
        all_cpus = os.sched_getaffinity(0)
        numa_node = None if args.no_numa_bind else ops.bind_host_to_device_numa(dev)   # pinned buffers local to the GPU
        host_in = [t.cpu().pin_memory() for t in (y_top, y_base, mu, std)]
        q_host = q_all.cpu().pin_memory()
        host_out = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32).pin_memory()
                    for k in want}
        nbytes = int(L.pic_host_pipeline_bytes(n, chunk))
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        tb = table.cpu()
        torch.cuda.synchronize()

        def host_step():
            rc = L.pic_slice_forward_host(host_in[0].data_ptr(), host_in[1].data_ptr(), host_in[2].data_ptr(),
                                          host_in[3].data_ptr(), 0.5, q_host.data_ptr(), None, tb.data_ptr(), 64,
                                          0.11, 1e-9, n, units, chunk, host_out["mask"].data_ptr(),
                                          host_out["y_hat"].data_ptr(), host_out["lik"].data_ptr(),
                                          host_out["idx"].data_ptr(), None, None, None, dbuf.data_ptr(), nbytes)
            if rc != 0:
                raise RuntimeError(f"pic_slice_forward_host rc={rc}")

        host_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()          # returns when the outputs are in host memory
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tdt = torch.tensor([dt], device=dev)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": total_elems * e2e_steps / dt / 1e9, "unit": "Gelem/s",
               "h2d_bytes_per_step": int(units * n * 16 + units * 4), "d2h_bytes_per_step": int(units * n * 16),
               "steps": e2e_steps, "value_scope": "whole job (all ranks' elements / slowest rank's wall time)",
               "bytes_scope": "per rank and step",
               "host_numa_node": numa_node,
               "api": "pic_slice_forward_host (C ABI, pinned host buffers, 3-slot copy/compute pipeline)"}
        if rank == 0:
            assert torch.equal(host_out["mask"][:per_slice], outs["mask"][:per_slice].cpu()), "host/device mismatch"
        os.sched_setaffinity(0, all_cpus)   # the CPU baseline below gets every host core again

    # ---------------- roofline of the dominant kernel: CUDA events around that kernel alone ----------------
    hbm, which = peaks()
    roof = None
    if rank == 0 or world > 1:
        other = None
        if name == "kodak_sweep" and args.launch == "per_step" and step_kind == 0:
            kern_ms, per_launch_elems = ms_per_step, units * n       # the step IS one launch of the kernel
        elif name == "kodak_sweep" and args.launch == "per_step":
            # plan 1: select kernel + tile-ordered apply kernel.  Time the dominant (apply) kernel alone
            # with the thresholds given, and the select kernel alone, with CUDA events on this stream.
            def ev_time(fn, reps=20):
                fn()
                torch.cuda.synchronize()
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(reps):
                    fn()
                a1.record()
                torch.cuda.synchronize()
                return a0.elapsed_time(a1) / reps
            thr0 = ops.select_threshold(std, units, q_all)
            kern_ms = ev_time(lambda: ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, thr_in=thr0,
                                                        want=want, out=outs))
            sel_ms = ev_time(lambda: ops.select_threshold(std, units, q_all))
            per_launch_elems = units * n
            other = {"select_kernel": "slice_fused_kernel<select-only>", "select_ms": sel_ms,
                     "select_GBps_of_std": units * n * 4 / (sel_ms * 1e-3) / 1e9,
                     "apply_share_of_step": kern_ms / ms_per_step}
        else:
            reps = 10
            if name == "first_train":
                fn = lambda: ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, noise=noise, want=want, out=outs)  # noqa: E731
                per_launch_elems = units * n
            elif name == "rem_latent":
                fn = lambda: ops.slice_forward(y_top, y_base, mu, std, units, q_all, None, want=want, out=outs)  # noqa: E731
                per_launch_elems = units * n
            elif name == "tile8192":
                thr0 = ops.select_threshold(std, units, q_all) if world == 1 else pdist.tiled_select_threshold(std, units, n, q_all, backend=backend)
                fn = lambda: ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, thr_in=thr0, want=want, out=outs)  # noqa: E731
                per_launch_elems = elems_per_rank
            else:
                fn = lambda: ops.slice_forward(view(y_top, 0), view(y_base, 0), view(mu, 0), view(std, 0), per_slice, q_slice, table, want=want, out={k: view(v, 0) for k, v in outs.items()})  # noqa: E731
                per_launch_elems = per_slice * n
            fn()
            torch.cuda.synchronize()
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record()
            for _ in range(reps):
                fn()
            k1.record()
            torch.cuda.synchronize()
            kern_ms = k0.elapsed_time(k1) / reps
        kbytes = {"first_train": 32, "tile8192": 32, "rem_latent": 28}.get(name, wl["bytes_per_elem"])
        bytes_per_launch = per_launch_elems * kbytes
        achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.isfile(tpath):
            try:
                traffic = json.load(open(tpath)).get(name)
            except Exception:
                traffic = None
        roof = {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                "traffic": traffic, "peak_source": which, "kernel": main_kernel,
                "algorithmic_bytes_per_elem": kbytes, "algorithmic_bytes_per_launch": bytes_per_launch,
                "launch_ms": kern_ms, "other_kernels": other,
                "whole_step_frac": (elems_per_rank * wl["bytes_per_elem"] / (ms_per_step * 1e-3) / 1e9) / hbm}
    if tile_comm is not None:
        tile_comm.close()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    cpu = None if args.no_cpu else cpu_reference_throughput(wl, args.cpu_seconds)
    line = {
        "metric": METRIC, "value": value, "unit": "Gelem/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "ms_per_step_by_rank": [round(t, 5) for t in rank_ms],
        "eager_ms_per_step": eager_ms,
        "config": {"workload": name, "desc": wl["desc"], "n_per_unit": n, "slices": slices,
                   "units_per_step_per_gpu": units, "elements_per_step": total_elems, "launch": args.launch,
                   "cuda_graph": use_graph,
                   "launches_per_step": launches[0], "outputs": list(want), "bytes_per_elem": wl["bytes_per_elem"],
                   "l2": f"inputs of one step = {elems_per_rank * 16 / 1e6:.0f} MB per GPU > 126 MB L2, no explicit flush",
                   "parallelism": (f"{world} rank(s), row-band tiles + NCCL histogram all-reduce"
                                   + (" issued by libpic_latent.so" if tile_comm is not None else " over torch.distributed")
                                   if name == "tile8192"
                                   else f"{world} rank(s), units sharded, no collective")},
        "roofline": roof, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches[0] * args.steps, "clocks": clocks,
    }
    os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="kodak_sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--launch", default="per_step", choices=["per_step", "per_slice"],
                    help="kodak_sweep: all (slice, q) units of a step in one launch, or one launch per slice index")
    ap.add_argument("--graph", type=int, default=1, help="1: timed steps replay a CUDA graph of one step (default; NCCL steps stay eager)")
    ap.add_argument("--torch-collectives", action="store_true",
                    help="tile8192, N > 1: carry the histogram all-reduces over torch.distributed instead of the library's own NCCL calls")
    ap.add_argument("--no-numa-bind", action="store_true", help="e2e: do not bind the process to the GPU's NUMA node")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--e2e-chunk", type=int, default=0, help="units per pipeline chunk of the host-buffer path (0 = auto)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_cuda(args, wl)


if __name__ == "__main__":
    main()
