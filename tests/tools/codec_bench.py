"""Throughput of the codec side (include/pic_codec.h) at the Kodak shape: 10 slice streams x 49 152 symbols per
level.  Reports the native coder (host threads, int32 buffers in place), the same through the reference's list
boundary (`.tolist()` per stream, entropy_models.py:229-236), and the pure-Python oracle on a small sample.
With a GPU: also symbols / indexes produced by the latent path on the device (one pinned D2H copy).
usage: python tests/tools/codec_bench.py [levels]"""
import os, sys, time
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pic_b200 as pic
from pic_b200 import codec

levels = int(sys.argv[1]) if len(sys.argv) > 1 else 4
slices, n = 10, 49152
gc = pic.GaussianConditional(None)
table = torch.exp(torch.linspace(np.log(0.11), np.log(256.0), 64))
gc.update(table.tolist())
tables = gc._tables()
rng = np.random.default_rng(0)
std = np.exp(rng.normal(-1.0, 1.2, size=(slices, n))).clip(1e-3, 300).astype(np.float32)
idx = np.searchsorted(table.numpy()[:-1], np.maximum(std, 0.11), side="left").astype(np.int32)
sym = np.rint(rng.normal(0, 1, size=(slices, n)) * std).astype(np.int32)
level = rng.integers(0, levels, size=(slices, n))


def med(fn, reps=5):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); r = fn(); ts.append(time.perf_counter() - t)
    return sorted(ts)[len(ts) // 2], r


per_level = [((sym * (level == l)).astype(np.int32), (idx * (level == l)).astype(np.int32)) for l in range(levels)]
t_enc, streams = med(lambda: [codec.encode_streams(s, i, tables) for s, i in per_level])
t_dec, back = med(lambda: [codec.decode_streams(st, i, tables) for st, (_, i) in zip(streams, per_level)])
assert all(np.array_equal(b.numpy(), s) for b, (s, _) in zip(back, per_level))
total = levels * slices * n
nbytes = sum(len(b) for st in streams for b in st)
print(f"native  encode {total / t_enc / 1e6:8.1f} Msym/s   decode {total / t_dec / 1e6:8.1f} Msym/s   "
      f"({levels} levels x {slices} streams x {n}; {nbytes} bytes = {8 * nbytes / (slices * n):.3f} bit/elem)")
lvl32 = level.astype(np.int32)
t_el, st_l = med(lambda: codec.encode_levels(sym, idx, lvl32, levels, tables))
assert st_l == streams
t_dl, back_l = med(lambda: codec.decode_levels(st_l, idx, lvl32, tables))
assert np.array_equal(back_l.numpy(), sym)
print(f"levels  encode {total / t_el / 1e6:8.1f} Msym/s   decode {total / t_dl / 1e6:8.1f} Msym/s   "
      f"(encode_levels / decode_levels on host arrays: selection inside the coder, {t_el * 1e3:.2f} / {t_dl * 1e3:.2f} ms)")
t1, _ = med(lambda: [codec.encode_streams(s, i, tables, threads=1) for s, i in per_level], reps=3)
print(f"native  encode, 1 thread {total / t1 / 1e6:8.1f} Msym/s")
c = codec.RansCoder()
lists = (gc._quantized_cdf.tolist(), gc._cdf_length.tolist(), gc._offset.tolist())
s0, i0 = per_level[0]
t_list, _ = med(lambda: [c.encode_with_indexes(torch.from_numpy(s0[k]).tolist(), torch.from_numpy(i0[k]).tolist(), *lists)
                         for k in range(slices)], reps=3)
print(f"native through the reference's list boundary (.tolist per stream) {slices * n / t_list / 1e6:8.1f} Msym/s")
import rans_oracle as ro
m = 20000
t_or, _ = med(lambda: ro.encode_with_indexes(s0[0, :m].tolist(), i0[0, :m].tolist(), *lists), reps=1)
print(f"pure-Python oracle encode ({m} symbols) {m / t_or / 1e6:8.3f} Msym/s")
if torch.cuda.is_available():
    dev = torch.device("cuda:0")
    ds, di, dl = (torch.from_numpy(a).to(dev) for a in (sym, idx, level.astype(np.int32)))
    gcd = pic.GaussianConditional(None).to(dev); gcd.update(table.tolist())
    def run():
        out = []
        for l in range(levels):
            d = (dl == l).to(torch.int32)
            out.append(gcd.compress(ds * d, di * d, already_quantize=True))
        return out
    run(); torch.cuda.synchronize()
    t_gpu, st2 = med(run)
    assert st2 == streams
    print(f"device tensors -> compress() ({levels} levels)  {t_gpu * 1e3:7.2f} ms  = {total / t_gpu / 1e6:8.1f} Msym/s")
    t_lv, st3 = med(lambda: codec.encode_levels(ds, di, dl, levels, gcd._tables()))
    assert st3 == streams
    print(f"device tensors -> encode_levels() (one D2H, {levels} levels)  {t_lv * 1e3:7.2f} ms  = {total / t_lv / 1e6:8.1f} Msym/s")
