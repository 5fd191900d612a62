"""Worker of tests/test_gpu_multi.py: launched with torch.distributed.run, one rank per GPU (NCCL).
Every rank holds a RAGGED band of each unit; the thresholds of the whole units come from the tiled select --
once with the collectives issued inside libpic_latent.so (pic_tiled_select_threshold), once carried by
torch.distributed -- and must equal the oracle's quantile of the full unit bit for bit on every rank; the local
masks / outputs must equal the oracle's slice restricted to the band."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
TESTS = os.path.dirname(HERE)
ROOT = os.path.dirname(TESTS)
for p in (ROOT, TESTS, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import pic_oracle as po  # noqa: E402  (test infrastructure: the checker)
from _common import LIK_ATOL, LIK_RTOL, scale_table, trained_like  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    import pic_b200
    from pic_b200 import distributed as pdist
    from pic_b200 import ops

    po.build()
    units, n = 7, 3 * 40000 + 7                     # not divisible by the rank count, not a multiple of 4
    rng = np.random.default_rng(2024)                # same data on every rank
    y_top, y_base, mu, std = trained_like(rng, (units, n))
    std[:, ::3] = np.round(std[:, ::3] * 8) / 8      # ties across the band boundaries
    std[1, :] = 0.5                                   # all-equal unit
    std[2, 11] = np.nan                               # NaN unit -> NaN threshold, all-false mask
    prs = [0.5, 5, 9.9999, 0, 10, 1e-4, 7.3]
    # ragged bands: rank r gets a share proportional to r + 1
    cuts = np.concatenate([[0], np.cumsum([(r + 1) for r in range(world)])]) * n // (world * (world + 1) // 2)
    cuts[-1] = n
    lo, hi = int(cuts[rank]), int(cuts[rank + 1])
    band = lambda a: torch.from_numpy(np.ascontiguousarray(a[:, lo:hi])).to(dev)  # noqa: E731
    q = ops.q01_tensor(prs, dev)
    comm = pdist.NcclTileComm(dev)
    std_l = band(std)
    thr_lib = pdist.tiled_select_threshold(std_l, units, n, q, comm=comm)
    thr_td = pdist.tiled_select_threshold(std_l, units, n, q)
    table = scale_table()
    ref = po.slice_forward(y_top, y_base, mu, std, prs, table)
    ok = True
    msgs = []

    def check(cond, what):
        nonlocal ok
        if not cond:
            ok = False
            msgs.append(what)

    check(np.array_equal(thr_lib.cpu().numpy(), ref["thr"], equal_nan=True), "library-issued NCCL thresholds != oracle")
    check(np.array_equal(thr_td.cpu().numpy(), ref["thr"], equal_nan=True), "torch.distributed thresholds != oracle")
    every = [torch.zeros_like(thr_lib) for _ in range(world)]
    dist.all_gather(every, thr_lib)
    check(all(np.array_equal(every[0].cpu().numpy(), e.cpu().numpy(), equal_nan=True) for e in every), "ranks disagree")
    out = pdist.tiled_slice_forward(band(y_top), band(y_base), band(mu), std_l, units, n, q, torch.from_numpy(table).to(dev),
                                    want=("mask", "y_hat", "lik", "idx"), comm=comm)
    torch.cuda.synchronize()
    for k in ("mask", "y_hat", "idx"):
        check(np.array_equal(out[k].cpu().numpy(), ref[k][:, lo:hi]), f"{k} of the band != oracle")
    lik, want = out["lik"].cpu().numpy().astype(np.float64), ref["lik"][:, lo:hi].astype(np.float64)
    both_nan = np.isnan(lik) & np.isnan(want)       # the NaN scale of unit 2 gives a NaN likelihood on both sides
    close = np.abs(lik - want) <= LIK_RTOL * np.abs(want) + LIK_ATOL
    check(bool(np.all(close | both_nan)), "likelihood of the band out of tolerance")
    # sampled protocol (C ABI 1d) on the same ragged, tie-heavy bands: exact whether or not it has to fall back
    thr_s = pdist.tiled_select_threshold(std_l, units, n, q, comm=comm, protocol="sampled")
    check(np.array_equal(thr_s.cpu().numpy(), ref["thr"], equal_nan=True), "sampled protocol (ragged bands) != oracle")
    # ... and on equal bands of larger iid units, where it must NOT fall back
    units2, n2 = 5, world * 262144
    rng2 = np.random.default_rng(77)
    std2 = trained_like(rng2, (units2, n2))[3]
    std2[3, 5] = np.nan
    prs2 = [0.5, 2.5, 5.0, 9.0, 7.0]
    b2 = n2 // world
    std2_l = torch.from_numpy(np.ascontiguousarray(std2[:, rank * b2:(rank + 1) * b2])).to(dev)
    before = comm.fallbacks
    thr2 = pdist.tiled_select_threshold(std2_l, units2, n2, ops.q01_tensor(prs2, dev), comm=comm, protocol="sampled")
    want2 = np.asarray([po.quantile(std2[u], np.float32(1.0 - prs2[u] * 0.1))[0] for u in range(units2)], np.float32)
    check(np.array_equal(thr2.cpu().numpy(), want2, equal_nan=True), f"sampled protocol (equal bands) != oracle: {thr2.cpu().numpy()} vs {want2}")
    check(comm.fallbacks == before, f"sampled protocol fell back on iid equal bands ({comm.fallbacks - before})")
    thr2r = pdist.tiled_select_threshold(std2_l, units2, n2, ops.q01_tensor(prs2, dev), comm=comm)
    check(np.array_equal(thr2r.cpu().numpy(), want2, equal_nan=True), "rounds protocol (equal bands) != oracle")
    # peer-memory transport of the sampled protocol (C ABI 1e): same thresholds, no NCCL call in the select itself
    comm.enable_p2p(max(n, n2), max(units, units2))
    for rep in range(3):                              # windows are reused: epochs have to line up call after call
        thr_p = pdist.tiled_select_threshold(std_l, units, n, q, comm=comm, protocol="p2p")
        check(np.array_equal(thr_p.cpu().numpy(), ref["thr"], equal_nan=True), f"p2p protocol (ragged bands, call {rep}) != oracle")
        before = comm.fallbacks
        thr2p = pdist.tiled_select_threshold(std2_l, units2, n2, ops.q01_tensor(prs2, dev), comm=comm, protocol="p2p")
        check(np.array_equal(thr2p.cpu().numpy(), want2, equal_nan=True), f"p2p protocol (equal bands, call {rep}) != oracle")
        check(comm.fallbacks == before, "p2p protocol fell back on iid equal bands")
    # bands of different content (a threshold inside the last band's value range puts most of the bracket's elements
    # there): the exchange slot holds up to 4 x a rank's fair share, so up to 4 ranks this must not fall back
    std3 = std2.copy()
    for r in range(world):
        std3[:, r * b2:(r + 1) * b2] *= np.float32(1.0 + 0.75 * r)
    std3_l = torch.from_numpy(np.ascontiguousarray(std3[:, rank * b2:(rank + 1) * b2])).to(dev)
    want3 = np.asarray([po.quantile(std3[u], np.float32(1.0 - prs2[u] * 0.1))[0] for u in range(units2)], np.float32)
    before = comm.fallbacks
    thr3 = pdist.tiled_select_threshold(std3_l, units2, n2, ops.q01_tensor(prs2, dev), comm=comm, protocol="p2p")
    check(np.array_equal(thr3.cpu().numpy(), want3, equal_nan=True), "p2p protocol (bands of different content) != oracle")
    if world <= 4:
        check(comm.fallbacks == before, f"p2p protocol fell back on bands of different content ({comm.fallbacks - before})")
    # asynchronous form, as a CUDA graph would replay it: status word read once at the end
    comm._p2p_status.zero_()
    outs_async = [comm.select_threshold(std2_l, units2, n2, ops.q01_tensor(prs2, dev), protocol="p2p", check_status=False)
                  for _ in range(20)]
    check(comm.p2p_status() == 0, f"p2p status word {comm.p2p_status():#x} after 20 back-to-back selects")
    check(all(np.array_equal(o.cpu().numpy(), want2, equal_nan=True) for o in outs_async), "back-to-back p2p selects != oracle")
    comm.close()
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if not ok:
        print(f"[rank {rank}] FAILED: " + "; ".join(msgs), flush=True)
    dist.destroy_process_group()
    if rank == 0 and int(flag.item()) == 1:
        print("TILED_TWO_RANK_OK", flush=True)
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
