"""TEST INFRASTRUCTURE: numpy stand-in for CudaTileBackend so that the all-reduce protocol of
pic_b200.distributed.tiled_select_threshold can run under gloo on CPU (world_size 2).
Mirrors the C ABI section (1b) step by step: begin / hist_round / advance / finish."""
import numpy as np
import torch

WORDS = 2056
SHIFT = (21, 10, 0)
BINS = (2048, 2048, 1024)


def f2key(x):
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).copy()
    b[b == 0x80000000] = 0
    neg = (b >> 31).astype(bool)
    return np.where(neg, ~b, b | np.uint32(0x80000000)).astype(np.uint32)


def key2f(k):
    k = np.asarray(k, dtype=np.uint32)
    pos = (k >> 31).astype(bool)
    return np.where(pos, k & np.uint32(0x7FFFFFFF), ~k).astype(np.uint32).view(np.float32)


class NumpyTileBackend:
    def __init__(self, std_local: np.ndarray, units: int):
        self.std = np.ascontiguousarray(std_local, dtype=np.float32).reshape(units, -1)
        self.keys = f2key(self.std)
        self.units = units
        self.min_above = torch.full((units,), -1, dtype=torch.int32)  # 0xffffffff

    def begin(self, n_total, q01):
        q = np.broadcast_to(np.asarray(q01.numpy() if isinstance(q01, torch.Tensor) else q01, np.float32), (self.units,))
        self.mode = np.where(q < 0, 1, np.where(q > 1, 0, 2))
        rank = (q * np.float32(n_total - 1)).astype(np.float32)
        self.lo = np.floor(rank).astype(np.int64)
        self.hi = np.ceil(rank).astype(np.int64)
        self.w = (rank - self.lo.astype(np.float32)).astype(np.float32)
        self.prefix = np.zeros(self.units, np.uint32)
        self.rank = self.lo.copy()
        self.below = np.zeros(self.units, np.int64)
        self.nan = np.zeros(self.units, bool)
        self.a = np.zeros(self.units, np.uint32)
        self.b = np.zeros(self.units, np.uint32)
        self.need_min = np.zeros(self.units, bool)

    def hist_round(self, r):
        hist = np.zeros((self.units, WORDS), np.int64)
        mn = np.full(self.units, 0xFFFFFFFF, np.uint64)
        for u in range(self.units):
            if self.mode[u] != 2:
                continue
            k = self.keys[u]
            if r == 0:
                hist[u, 2048] = int(np.isnan(self.std[u]).any())
                sel = k
            else:
                up = SHIFT[r] + (11 if r == 1 else 10)
                hb, want = k >> np.uint32(up), self.prefix[u] >> np.uint32(up)
                sel = k[hb == want]
                if r == 2 and (hb > want).any():
                    mn[u] = k[hb > want].min()
            digits = (sel >> np.uint32(SHIFT[r])) & np.uint32(BINS[r] - 1)
            hist[u, :BINS[r]] = np.bincount(digits, minlength=BINS[r])
        self.min_above = torch.from_numpy(mn.astype(np.uint32).view(np.int32).copy())
        return torch.from_numpy(hist.astype(np.int32).reshape(-1))

    def advance(self, hist, r):
        h = hist.numpy().reshape(self.units, WORDS).astype(np.int64)
        for u in range(self.units):
            if self.mode[u] != 2:
                continue
            c = np.cumsum(h[u, :BINS[r]])
            b = int(np.searchsorted(c, self.rank[u], side="right"))
            below = int(c[b - 1]) if b > 0 else 0
            self.prefix[u] |= np.uint32(b << SHIFT[r])
            self.rank[u] -= below
            self.below[u] += below
            if r == 0:
                self.nan[u] = h[u, 2048] != 0
            if r == 2:
                self.a[u] = self.prefix[u]
                cnt = int(h[u, b])
                if self.hi[u] < self.below[u] + cnt:
                    self.b[u] = self.a[u]
                else:
                    nz = np.nonzero(h[u, b + 1:BINS[2]])[0]
                    if nz.size:
                        self.b[u] = (self.prefix[u] & ~np.uint32(1023)) | np.uint32(b + 1 + nz[0])
                    else:
                        self.need_min[u] = True

    def min_above_keys(self):
        return self.min_above

    def finish(self, min_above):
        mn = min_above.numpy().view(np.uint32)
        thr = np.empty(self.units, np.float32)
        for u in range(self.units):
            if self.mode[u] == 1:
                thr[u] = -np.inf
            elif self.mode[u] == 0:
                thr[u] = np.inf
            elif self.nan[u]:
                thr[u] = np.nan
            else:
                a = key2f(self.a[u])
                b = key2f(mn[u] if self.need_min[u] else self.b[u])
                a, b, w = np.float32(a), np.float32(b), self.w[u]
                d = np.float32(b - a)
                if abs(w) < 0.5:
                    thr[u] = np.float32(np.float64(w) * np.float64(d) + np.float64(a))
                else:
                    thr[u] = np.float32(np.float64(b) - np.float64(d) * np.float64(np.float32(np.float32(1) - w)))
        return torch.from_numpy(thr)
