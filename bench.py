#!/usr/bin/env python
"""bench.py -- mask+quantize+likelihood throughput of the PIC latent hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # reference CPU path (oracle port) on host cores
    torchrun --nproc-per-node N bench.py --gpus N ...        # N>1: one rank per GPU, units sharded, no collective

Default workload = BASELINE config[1]: Kodak-shape 768x512 latents ([1,32,32,48] => n=49152 per
unit), 10 progressive slices x 101-point quality sweep (pr = 0, 0.1, .. 10) = 1010 units per step.
A step launches the fused slice kernel once per slice index (10 launches of 101 units: slices are
sequential in the codec, the quality sweep is the batch).  Inputs of a step are 794 MB (> 126 MB L2),
so consecutive steps stream from HBM without an explicit L2 flush.

One JSON line is printed by rank 0; see DESIGN.md "Measurement" for every key.
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
    # name: (n_per_unit, slices, q-points, description)
    "kodak_sweep": dict(n=32 * 32 * 48, slices=10, prs=[10.0 * k / 100 for k in range(101)],
                        desc="Kodak 768x512 batch 1, 10 progressive slices x 101-point quality sweep"),
    "first_train": dict(n=32 * 16 * 16, slices=10, prs=None, batch=256,
                        desc="256 crops of 256x256, 10 slices, random q per image (refine/rems sampling)"),
    "tile8192": dict(n=32 * 512 * 512, slices=10, prs=[1.0],
                     desc="single 8192x8192 image, 10 slices, q=1"),
}
BYTES_PER_ELEM = 32  # y_top 4 + y_base 4 + mu 4 + std 4 (read once: tile kept in smem) + mask, y_hat, lik, idx 16


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(path):
        try:
            with open(path) as f:
                return float(json.load(f)["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None
        self.thread = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return
        self.thread = threading.Thread(target=self._read, daemon=True)
        self.thread.start()

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0: float, t1: float):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                clk, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            smax.append(mx)
            if t0 - 0.05 <= ts <= t1 + 0.15:
                sm.append(clk)
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:  # timed region shorter than the sampling period: use all samples
            sm = [float(l.split(",")[0]) for _, l in self.rows if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


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
def cpu_reference_throughput(wl, seconds_budget, threads=None, sample_units=None):
    """Times the oracle port (oracle/pic_oracle.c, OpenMP over units) on a bounded sample."""
    import pic_oracle as po

    po.build()
    if threads:
        po.set_num_threads(threads)
    cores = po.num_threads()
    n = wl["n"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    if sample_units is None:
        sample_units = max(1, min(per_slice * wl["slices"], max(cores * 2, (4 << 20) // n)))
    prs_all = unit_prs(wl, per_slice, 4321)
    prs = [prs_all[u % len(prs_all)] for u in range(sample_units)]
    y_top, y_base, mu, std = make_host_inputs(n, sample_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    want = ("mask", "y_hat", "lik", "idx")
    po.slice_forward(y_top, y_base, mu, std, prs, table, want=want)  # warm-up
    times = []
    t_start = time.perf_counter()
    while True:
        t0 = time.perf_counter()
        po.slice_forward(y_top, y_base, mu, std, prs, table, want=want)
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_start > seconds_budget or len(times) >= 50:
            break
    med = statistics.median(times)
    gelem = sample_units * n / med / 1e9
    return {"value": gelem, "unit": "Gelem/s", "cores": cores, "kind": "port",
            "sample": f"{sample_units} units x {n} elem, median of {len(times)} passes of oracle/pic_oracle.c "
                      f"(qsort quantile + erfc + 63-step index), {cores} OpenMP threads"}, med, sample_units


def run_reference(args, wl):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import pic_oracle as po

    po.build()
    cores = po.num_threads()
    n = wl["n"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    sample_units = max(1, min(per_slice * wl["slices"], max(cores * 2, (4 << 20) // n)))
    prs_all = unit_prs(wl, per_slice, 4321)
    prs = [prs_all[u % len(prs_all)] for u in range(sample_units)]
    y_top, y_base, mu, std = make_host_inputs(n, sample_units, 99)
    table = np.load(os.path.join(ROOT, "tests", "golden", "scale_table.npy"))
    want = ("mask", "y_hat", "lik", "idx")
    for _ in range(args.warmup):
        po.slice_forward(y_top, y_base, mu, std, prs, table, want=want)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        po.slice_forward(y_top, y_base, mu, std, prs, table, want=want)
    dt = time.perf_counter() - t0
    value = args.steps * sample_units * n / dt / 1e9
    sample = (f"each step = {sample_units} units x {n} elem of the {args.workload} workload through "
              f"oracle/pic_oracle.c (C port of the reference's torch CPU path; the reference is Python and cannot "
              f"travel to this box), {cores} OpenMP threads")
    line = {"impl": "reference", "metric": "mask+quantize+likelihood throughput", "value": value, "unit": "Gelem/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "desc": wl["desc"], "n_per_unit": n,
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
    from pic_b200 import ops

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback; use --impl reference for the CPU path)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = pic_b200.lib()

    n, slices = wl["n"], wl["slices"]
    per_slice = len(wl["prs"]) if wl["prs"] is not None else wl.get("batch", 1)
    units = per_slice * slices                     # per rank (weak scaling: every rank runs a full replica)
    prs = unit_prs(wl, per_slice, 4321 + rank)
    seed = 1234 + 1000 * 2 + rank
    y_top, y_base, mu, std = make_device_inputs(torch, n, units, seed, dev)
    q_slice = ops.q01_tensor(prs, dev)
    table = pic_b200.get_scale_table().to(dev)
    want = ("mask", "y_hat", "lik", "idx")
    outs = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32, device=dev) for k in want}
    ws_bytes = int(L.pic_workspace_bytes(n, per_slice))
    fused = n <= int(L.pic_fused_max_elems())

    def view(t, s):
        return t[s * per_slice:(s + 1) * per_slice]

    launches_per_step = [0]

    q_all = torch.cat([q_slice] * slices).contiguous()

    def step():
        cnt = 0
        if args.launch == "per_step":
            ops.slice_forward(y_top, y_base, mu, std, units, q_all, table, want=want, out=outs)
            cnt = 1 if fused else 9
        else:
            for s in range(slices):
                o = {k: view(v, s) for k, v in outs.items()}
                ops.slice_forward(view(y_top, s), view(y_base, s), view(mu, s), view(std, s), per_slice, q_slice,
                                  table, want=want, out=o)
                cnt += 1 if fused else 9  # rounds path: begin + 3x(hist, advance) + finish + apply
        launches_per_step[0] = cnt

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # repeat the K timed steps a few times so that the clock sampler sees the load; report the best-of
    # is NOT done: the K steps are timed exactly once, the extra repetitions only feed the sampler.
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.perf_counter()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    t_wall1 = time.perf_counter()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms.item())
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms / args.steps
    total_elems = units * n * world
    value = total_elems / (ms_per_step * 1e-3) / 1e9

    # ---------------- end-to-end: host buffers through the C ABI (H2D + kernels + D2H timed) -------------
    e2e = None
    if not args.no_e2e:
        e2e_units = units
        chunk = max(1, min(per_slice, (48 << 20) // (n * 4)))
        host_in = [t.cpu().pin_memory() for t in (y_top, y_base, mu, std)]
        q_host = torch.cat([q_slice.cpu()] * slices).pin_memory()
        host_out = {k: torch.empty((units, n), dtype=torch.int32 if k == "idx" else torch.float32).pin_memory()
                    for k in want}
        nbytes = int(L.pic_host_pipeline_bytes(n, chunk))
        dbuf = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        tb = table.cpu()
        torch.cuda.synchronize()

        def host_step():
            rc = L.pic_slice_forward_host(host_in[0].data_ptr(), host_in[1].data_ptr(), host_in[2].data_ptr(),
                                          host_in[3].data_ptr(), 0.5, q_host.data_ptr(), None, tb.data_ptr(), 64,
                                          0.11, 1e-9, n, e2e_units, chunk, host_out["mask"].data_ptr(),
                                          host_out["y_hat"].data_ptr(), host_out["lik"].data_ptr(),
                                          host_out["idx"].data_ptr(), None, None, None, dbuf.data_ptr(), nbytes)
            if rc != 0:
                raise RuntimeError(f"pic_slice_forward_host rc={rc}")

        host_step()
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            host_step()          # blocks until the outputs are in host memory
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            tdt = torch.tensor([dt], device=dev)
            dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
            dt = float(tdt.item())
        e2e = {"value": e2e_units * n * world * e2e_steps / dt / 1e9, "unit": "Gelem/s",
               "h2d_bytes_per_step": int(e2e_units * n * 16 + e2e_units * 4),
               "d2h_bytes_per_step": int(e2e_units * n * 16), "steps": e2e_steps,
               "api": "pic_slice_forward_host (C ABI, pinned host buffers, 3-slot copy/compute pipeline)"}
        # cheap sanity: the host path and the device path agree
        if rank == 0:
            assert torch.equal(host_out["mask"][:per_slice], outs["mask"][:per_slice].cpu()), "host/device mismatch"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    hbm, which = peaks()
    launches = launches_per_step[0] * args.steps
    n_launch = 1 if args.launch == "per_step" else slices
    kern_ms = ms_per_step / n_launch  # fused path: the step is n_launch back-to-back launches of one kernel
    bytes_per_launch = units // n_launch * n * BYTES_PER_ELEM
    achieved = bytes_per_launch / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.isfile(tpath):
        try:
            traffic = json.load(open(tpath)).get(args.workload)
        except Exception:
            traffic = None
    cpu = None
    if not args.no_cpu:
        cpu, _, _ = cpu_reference_throughput(wl, args.cpu_seconds)
    line = {
        "metric": "mask+quantize+likelihood throughput", "value": value, "unit": "Gelem/s", "n_gpus": world,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": wl["desc"], "n_per_unit": n, "slices": slices,
                   "units_per_launch": units if args.launch == "per_step" else per_slice, "launch": args.launch, "units_per_step_per_gpu": units, "launches_per_step": launches_per_step[0],
                   "outputs": list(want), "bytes_per_elem": BYTES_PER_ELEM,
                   "l2": f"inputs of one step = {units * n * 16 / 1e6:.0f} MB > 126 MB L2, no explicit flush",
                   "parallelism": f"units sharded, {world} rank(s), no collective"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm, "unit": "GB/s", "frac": achieved / hbm,
                     "traffic": traffic, "peak_source": which, "kernel": "slice_fused_kernel" if fused else "slice_apply_kernel",
                     "algorithmic_bytes_per_launch": bytes_per_launch, "launch_ms": kern_ms},
        "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="kodak_sweep", choices=sorted(WORKLOADS))
    ap.add_argument("--launch", default="per_step", choices=["per_step", "per_slice"],
                    help="per_step: all (slice, q) units of a step in one launch; per_slice: one launch per slice index")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_cuda(args, wl)


if __name__ == "__main__":
    main()
