"""Multi-GPU forms of the path.  One process per GPU, torch.distributed for plumbing.

* Batch sharding (BASELINE configs 3/4): units are independent -- ``shard_range`` gives each
  rank its slice of the image batch; the path needs NO data-path collective.
* Spatial tiling (config 5, one huge image): every rank holds a band of each unit's latent.
  Everything is elementwise except the per-unit threshold, which needs global order
  statistics: per radix round (11/11/10 key bits) each rank histograms its tile
  (pic_hist_round), the histograms are all-reduced (sum, 2056 uint32 words per unit -- a
  latency-bound NCCL message over NVLink/NVSwitch), every rank advances the same state
  (pic_select_advance); one more all-reduce (min) supplies the successor key when it lies
  outside the final bucket.  All ranks then hold bit-identical thresholds and apply them
  locally (pic_slice_forward with thr_in), so masks equal the single-GPU / CPU result.

The local kernels sit behind a small backend interface so the collective protocol can be
exercised on CPU with gloo in the test-suite (tests inject a numpy backend; the product
backend is CUDA-only).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist

from . import ops
from ._lib import check, lib

_SIGN = -(2 ** 31)


def shard_range(total_units: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the unit batch owned by `rank` (sizes differ by <= 1)."""
    base, rem = divmod(total_units, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class CudaTileBackend:
    """Local steps of the tiled select on this rank's CUDA tile (C ABI section 1b)."""

    def __init__(self, std_local: torch.Tensor, units: int):
        self.std = ops._require(std_local, "std_local")
        self.units = units
        self.n_local = self.std.numel() // units if units else 0
        dev = self.std.device
        self.state = torch.empty(int(lib().pic_select_state_bytes(units)), dtype=torch.uint8, device=dev)
        self.words = int(lib().pic_hist_words())
        self.hist = torch.empty(units * self.words, dtype=torch.int32, device=dev)
        self.min_above = torch.empty(units, dtype=torch.int32, device=dev)

    def begin(self, n_total: int, q01) -> None:
        q, qt = ops._q_args(q01, self.units, self.std.device)
        check(lib().pic_select_begin(self.state.data_ptr(), n_total, self.units, q, ops._ptr(qt), ops._stream()),
              "pic_select_begin")

    def hist_round(self, rnd: int) -> torch.Tensor:
        check(lib().pic_hist_round(self.std.data_ptr(), self.n_local, self.units, rnd, self.state.data_ptr(),
                                   self.hist.data_ptr(), self.min_above.data_ptr(), ops._stream()), "pic_hist_round")
        return self.hist

    def advance(self, hist: torch.Tensor, rnd: int) -> None:
        check(lib().pic_select_advance(self.state.data_ptr(), hist.data_ptr(), self.units, rnd, ops._stream()),
              "pic_select_advance")

    def min_above_keys(self) -> torch.Tensor:
        return self.min_above

    def finish(self, min_above: torch.Tensor) -> torch.Tensor:
        thr = torch.empty(self.units, dtype=torch.float32, device=self.std.device)
        check(lib().pic_select_finish(self.state.data_ptr(), min_above.data_ptr(), self.units, thr.data_ptr(),
                                      None, None, ops._stream()), "pic_select_finish")
        return thr


class NcclTileComm:
    """NCCL communicator owned by libpic_latent.so for `pic_tiled_select_threshold` (C ABI 1c): the whole tiled
    select becomes ONE host call that enqueues kernels and collectives on the current stream.  The 128-byte
    unique id travels over torch.distributed (`group`), which must already be initialised."""

    def __init__(self, device: torch.device, group=None):
        import ctypes

        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * 128)()
            check(lib().pic_dist_unique_id(buf), "pic_dist_unique_id")
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        ident = ident.to(device) if dist.get_backend(group) == "nccl" else ident
        dist.broadcast(ident, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        raw = bytes(ident.cpu().tolist())
        handle = ctypes.c_void_p()
        with torch.cuda.device(device):
            check(lib().pic_dist_comm_init(raw, rank, world, ctypes.byref(handle)), "pic_dist_comm_init")
        self.handle, self.device, self.world = handle, device, world
        self._ws = None
        self.fallbacks = 0            # sampled / p2p selects that had to run the histogram rounds
        self.rank = rank
        self._p2p = None
        self._p2p_status = None

    def enable_p2p(self, n_total_max: int, units_max: int) -> None:
        """Collective over the communicator: allocate this rank's peer-memory window (sized for the largest tiled unit /
        slice count it will serve), exchange the CUDA IPC handles and map the peers' windows (C ABI 1e).  Single node."""
        import ctypes

        if self._p2p is not None:
            return
        a, b = ctypes.c_size_t(0), ctypes.c_size_t(0)
        check(lib().pic_dist_p2p_region_bytes(n_total_max, units_max, self.world, ctypes.byref(a), ctypes.byref(b)),
              "pic_dist_p2p_region_bytes")
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            check(lib().pic_dist_p2p_init(self.handle, self.rank, a.value, b.value, ctypes.byref(handle)), "pic_dist_p2p_init")
        self._p2p = handle
        self._p2p_status = torch.zeros(1, dtype=torch.int32, device=self.device)

    def p2p_status(self) -> int:
        """Status word of the p2p selects issued since it was last cleared (synchronises): low 16 bits = units whose
        bracket missed, higher bits = exchange waits that timed out.  Non-zero: those thresholds are not valid."""
        return int(self._p2p_status.item()) & 0xFFFFFFFF

    def select_threshold(self, std_local: torch.Tensor, units: int, n_total: int, q01, protocol: str = "rounds",
                         check_status: bool = True) -> torch.Tensor:
        """protocol "rounds": three histogram all-reduces + a min all-reduce, fully asynchronous (graph-capturable);
        "sampled": two all-gathers and one pass over the band, one stream synchronisation per call (C ABI 1d);
        "p2p": the sampled protocol over peer-memory windows (enable_p2p first; C ABI 1e) -- with check_status=False
        nothing synchronises (graph-capturable) and the caller reads p2p_status() at its next synchronisation point."""
        import ctypes

        if protocol not in ("rounds", "sampled", "p2p"):
            raise ValueError(f"unknown tiled-select protocol {protocol!r}")
        if protocol == "p2p" and self._p2p is None:
            raise RuntimeError("NcclTileComm.enable_p2p(n_total_max, units_max) has to run (on every rank) first")
        std_local = ops._require(std_local, "std_local")
        n_local = std_local.numel() // units
        q, qt = ops._q_args(q01, units, std_local.device)
        thr = torch.empty(units, dtype=torch.float32, device=std_local.device)
        if protocol == "p2p":
            need = int(lib().pic_tiled_sampled_workspace_bytes(n_local, n_total, units, self.world))
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=std_local.device)
            if check_status:
                self._p2p_status.zero_()
            check(lib().pic_tiled_select_threshold_p2p(ops._ptr(std_local), n_local, n_total, units, q, ops._ptr(qt), ops._ptr(thr),
                                                       ops._ptr(self._ws), self._ws.numel(), self._p2p,
                                                       ops._ptr(self._p2p_status), ops._stream()), "pic_tiled_select_threshold_p2p")
            if check_status:
                status = self.p2p_status()
                if status >> 16:
                    raise RuntimeError("peer-memory exchange timed out: a rank of the tiled select never arrived")
                if status:
                    self.fallbacks += 1
                    return self.select_threshold(std_local, units, n_total, q01, protocol="rounds")
            return thr
        if protocol == "sampled":
            need = int(lib().pic_tiled_sampled_workspace_bytes(n_local, n_total, units, self.world))
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty(need, dtype=torch.uint8, device=std_local.device)
            fb = ctypes.c_int(0)
            check(lib().pic_tiled_select_threshold_sampled(ops._ptr(std_local), n_local, n_total, units, q, ops._ptr(qt),
                                                           ops._ptr(thr), ops._ptr(self._ws), self._ws.numel(), self.handle,
                                                           ops._stream(), ctypes.byref(fb)), "pic_tiled_select_threshold_sampled")
            self.fallbacks += fb.value
            return thr
        need = int(lib().pic_tiled_workspace_bytes(units))
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, dtype=torch.uint8, device=std_local.device)
        check(lib().pic_tiled_select_threshold(ops._ptr(std_local), n_local, n_total, units, q, ops._ptr(qt), ops._ptr(thr),
                                               ops._ptr(self._ws), self._ws.numel(), self.handle, ops._stream()),
              "pic_tiled_select_threshold")
        return thr

    def close(self) -> None:
        if self._p2p is not None:
            torch.cuda.synchronize(self.device)
            with torch.cuda.device(self.device):
                check(lib().pic_dist_p2p_destroy(self._p2p), "pic_dist_p2p_destroy")
            self._p2p = None
        if self.handle:
            torch.cuda.synchronize(self.device)
            check(lib().pic_dist_comm_destroy(self.handle), "pic_dist_comm_destroy")
            self.handle = None


def allreduce_min_u32(keys_i32: torch.Tensor, group=None) -> torch.Tensor:
    """MIN over ranks of uint32 keys stored in an int32 tensor (torch has no uint32 collectives):
    flipping the top bit maps unsigned order onto signed order."""
    flipped = keys_i32 ^ _SIGN
    dist.all_reduce(flipped, op=dist.ReduceOp.MIN, group=group)
    return flipped ^ _SIGN


def tiled_select_threshold(std_local: Optional[torch.Tensor], units: int, n_total: int, q01, group=None,
                           backend=None, comm: Optional[NcclTileComm] = None, protocol: str = "rounds") -> torch.Tensor:
    """Global per-unit quantile threshold of units whose elements are spread over the ranks of
    `group`.  Returns thr [units], bit-identical on every rank.  `comm` (NcclTileComm): the collectives are
    issued inside the library (one host call); otherwise torch.distributed carries them round by round."""
    if comm is not None:
        return comm.select_threshold(std_local, units, n_total, q01, protocol=protocol)
    be = backend if backend is not None else CudaTileBackend(std_local, units)
    be.begin(n_total, q01)
    for rnd in range(3):
        hist = be.hist_round(rnd)
        dist.all_reduce(hist, op=dist.ReduceOp.SUM, group=group)  # counts < 2^24: int32 view is exact
        be.advance(hist, rnd)
    mn = allreduce_min_u32(be.min_above_keys(), group)
    return be.finish(mn)


def tiled_slice_forward(y_top, y_base, mu, std, units: int, n_total: int, pr, scale_table=None, noise=None,
                        group=None, scale_bound: float = 0.11, lik_bound: float = 1e-9,
                        want=("mask", "y_hat", "lik"), comm: Optional[NcclTileComm] = None, protocol: str = "rounds") -> dict:
    """One progressive slice of spatially tiled units: all-reduced threshold + local apply."""
    q01 = pr if isinstance(pr, torch.Tensor) else ops.pr_to_q01(pr)
    mode_scalar = None if isinstance(q01, torch.Tensor) else q01
    if mode_scalar is not None and not (0.0 <= mode_scalar <= 1.0):
        # ones / zeros short-circuit: no threshold, no collective
        return ops.slice_forward(y_top, y_base, mu, std, units, q01, scale_table, noise=noise,
                                 scale_bound=scale_bound, lik_bound=lik_bound, want=want)
    thr = tiled_select_threshold(std, units, n_total, q01, group, comm=comm, protocol=protocol)
    res = ops.slice_forward(y_top, y_base, mu, std, units, q01, scale_table, noise=noise, thr_in=thr,
                            scale_bound=scale_bound, lik_bound=lik_bound, want=want)
    res["thr"] = thr
    return res
