"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz by running the UNMODIFIED reference
functions (through oracle/ref_shim.py) on seeded inputs, in the build container.

    python oracle/gen_golden.py            # rewrites tests/golden/

The committed fixtures are what pins the C oracle (tests/test_oracle_golden.py) and, on the
GPU box (where /root/reference does not exist), the CUDA path (tests/test_gpu_parity.py).
Torch used for generation is recorded in tests/golden/MANIFEST.json.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


# ----------------------------------------------------------------------------- inputs
def trained_like(rng, shape):
    """SURVEY 8(d) 'S-trained-like' synthetic latents."""
    std = np.exp(rng.normal(-1.0, 1.2, size=shape)).clip(1e-3, 300.0)
    flip = rng.random(size=shape) < 0.02
    std = np.where(flip, -std, std).astype(np.float32)
    mu = rng.normal(0, 1, size=shape).astype(np.float32)
    y_base = rng.normal(0, 2, size=shape).astype(np.float32)
    y_top = (y_base + mu + np.abs(std) * rng.normal(0, 1, size=shape)).astype(np.float32)
    return y_top, y_base, mu, std


def model_like(rng, shape):
    """random-init-model-like: std in +-0.05, ~half negative (SURVEY 8(d) 'S-model')."""
    std = rng.normal(0, 0.02, size=shape).astype(np.float32)
    mu = rng.normal(0, 0.05, size=shape).astype(np.float32)
    y_base = rng.normal(0, 0.3, size=shape).astype(np.float32)
    y_top = (y_base + rng.normal(0, 0.8, size=shape)).astype(np.float32)
    return y_top, y_base, mu, std


def hashed_std(n: int, seed: int) -> np.ndarray:
    """Platform-independent large input: exact integer hash -> exact f32.  Regenerated (not
    stored) by the tests; values in [-0.5, 0.5) on a 2^-24 grid so ties are frequent."""
    i = np.arange(n, dtype=np.uint64)
    k = (i * np.uint64(2654435761) + np.uint64(seed) * np.uint64(40503)) & np.uint64(0xFFFFFFFF)
    k ^= k >> np.uint64(15)
    k = (k * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    k ^= k >> np.uint64(13)
    return ((k >> np.uint64(8)).astype(np.float32) / np.float32(1 << 24)) - np.float32(0.5)


def t(a):
    return torch.from_numpy(np.ascontiguousarray(a))


PR_LIST = [0, 1e-4, 0.5, 0.75, 1, 2.5, 5, 7.3, 9.9999, 10, 11]


# ----------------------------------------------------------------------------- generators
def gen_masks(ref):
    masking = ref.ChannelMask("point-based-std")
    rng = np.random.default_rng(1001)
    cases = {}
    shapes = {"n512": (2, 32, 4, 4), "n1120": (3, 32, 5, 7), "n8192": (1, 32, 16, 16)}
    for name, shape in shapes.items():
        _, _, _, std = trained_like(rng, shape)
        cases[f"trained_{name}"] = std
    cases["modellike_n2048"] = model_like(rng, (2, 32, 8, 8))[3]
    ties = np.round(rng.normal(0, 1, size=(2, 32, 8, 8)) * 64 / 16) * (16 / 64)
    cases["ties_n2048"] = ties.astype(np.float32)
    cases["allequal_n512"] = np.full((1, 32, 4, 4), 0.25, dtype=np.float32)
    z = rng.normal(0, 1, size=(2, 32, 4, 4)).astype(np.float32)
    z[np.abs(z) < 0.7] = 0.0
    z[0, :8][z[0, :8] == 0] = -0.0
    z[1, :4] = np.float32(1e-42) * rng.integers(-5, 5, size=z[1, :4].shape).astype(np.float32)
    cases["zeros_denormals_n512"] = z
    w_nan = trained_like(rng, (2, 32, 4, 4))[3]
    w_nan[1, 3, 2, 1] = np.nan
    cases["nan_n512"] = w_nan
    inf = trained_like(rng, (1, 32, 4, 4))[3]
    inf[0, 0, 0, 0] = np.inf
    inf[0, 1, 0, 0] = -np.inf
    cases["inf_n512"] = inf
    out = {}
    for name, std in cases.items():
        out[f"{name}/std"] = std
        for pr in PR_LIST:
            mask = masking(t(std), pr=pr)
            out[f"{name}/mask/pr={pr!r}"] = np.packbits(mask.numpy().astype(np.uint8).ravel())
            if 0 < pr < 10:
                thr = [torch.quantile(t(std)[j].ravel(), 1.0 - pr * 0.1).item() for j in range(std.shape[0])]
                out[f"{name}/thr/pr={pr!r}"] = np.asarray(thr, dtype=np.float32)
    # ProgMask: list of 10 per-slice [1,32,h,w] blocks (layers/channel_mask.py:18-49)
    blocks = [trained_like(rng, (1, 32, 4, 6))[3] for _ in range(10)]
    out["progmask/std"] = np.stack(blocks)
    for pr in PR_LIST:
        pm = masking.ProgMask([t(b) for b in blocks], pr)
        out[f"progmask/mask/pr={pr!r}"] = np.packbits(pm.numpy().astype(np.uint8).ravel())
        out[f"progmask/shape/pr={pr!r}"] = np.asarray(pm.shape)
    # ravel=True variant
    rv = trained_like(rng, (3, 700))[3]
    out["ravel/std"] = rv
    for pr in (0.5, 5):
        try:
            mk = masking(t(rv), pr=pr, ravel=True)
            out[f"ravel/mask/pr={pr!r}"] = mk.numpy()
        except Exception as e:  # reference bug: `scale[j,:,:,:]` is not reached when ravel
            out[f"ravel/error/pr={pr!r}"] = np.frombuffer(repr(e).encode(), dtype=np.uint8)
    return out


def gen_large_quantiles():
    """thr / order statistics only, on regenerable hashed inputs (C5-size and torch's max)."""
    out = {}
    for n, seed in ((8388608, 7), (1 << 24, 3), (1000003, 11), (49152, 5)):
        x = hashed_std(n, seed)
        xs = np.sort(x)
        for pr in (0.5, 1, 2.5, 5, 9.9999, 1e-4):
            q = 1.0 - pr * 0.1
            thr = torch.quantile(t(x), q).item()
            q32 = np.float32(q)
            rank = np.float32(q32 * np.float32(n - 1))
            lo, hi = int(np.floor(rank)), int(np.ceil(rank))
            cnt = int((x >= np.float32(thr)).sum())
            out[f"n={n}/seed={seed}/pr={pr!r}"] = np.asarray(
                [np.float32(thr), xs[lo], xs[hi], np.float32(cnt % 65536), np.float32(cnt // 65536)],
                dtype=np.float32)
    return out


def gen_slices(ref, gc):
    masking = ref.ChannelMask("point-based-std")
    rng = np.random.default_rng(2002)
    out = {}
    specs = [("trained_n512", trained_like, (2, 32, 4, 4)), ("trained_n2048", trained_like, (1, 32, 8, 8)),
             ("model_n2048", model_like, (2, 32, 8, 8)), ("trained_n3072", trained_like, (1, 32, 8, 12))]
    for name, fn, shape in specs:
        y_top, y_base, mu, std = fn(rng, shape)
        # exercise round-half-even and |d|=0.5 cut points
        d = y_top - y_base - mu
        y_top.ravel()[:8] = (y_base.ravel()[:8] + mu.ravel()[:8]
                             + np.asarray([0.5, 1.5, 2.5, -0.5, -1.5, 0.0, 3.5, -2.5], np.float32))
        del d
        for k, v in dict(y_top=y_top, y_base=y_base, mu=mu, std=std).items():
            out[f"{name}/{k}"] = v
        for pr in (0, 0.5, 1, 5, 7.3, 10):
            for training in (False, True):
                seed = 77 + int(pr * 10)
                if training:
                    torch.manual_seed(seed)
                    noise = torch.empty(shape).uniform_(-0.5, 0.5)
                    out[f"{name}/noise/pr={pr!r}"] = noise.numpy()
                yt, yb, m, s = (t(a).requires_grad_(True) for a in (y_top, y_base, mu, std))
                r = ref_shim.reference_slice_forward(ref, gc, masking, yt, yb, m, s, pr, training=training,
                                                     noise_seed=seed if training else None)
                tag = f"{name}/{'train' if training else 'eval'}/pr={pr!r}"
                if training:  # sanity: the reference really drew the same noise
                    y_m = (yt - yb - m) * r["mask"]
                    assert torch.equal((r["outputs"] - y_m).detach(), (y_m + noise - y_m).detach())
                out[f"{tag}/mask"] = np.packbits(r["mask"].numpy().astype(np.uint8).ravel())
                out[f"{tag}/lik"] = r["lik"].detach().numpy()
                out[f"{tag}/y_hat"] = r["y_hat"].detach().numpy()
                out[f"{tag}/outputs"] = r["outputs"].detach().numpy()
                out[f"{tag}/idx"] = r["idx"].numpy().astype(np.int8)
                out[f"{tag}/symbols"] = r["symbols"].numpy().astype(np.int32)
                out[f"{tag}/logsum"] = np.asarray(torch.log(r["lik"].detach()).sum().item(), np.float32)
                # backward with random cotangents
                g = np.random.default_rng(seed)
                g_lik = t(g.normal(0, 1, size=shape).astype(np.float32))
                g_y = t(g.normal(0, 1, size=shape).astype(np.float32))
                grads = torch.autograd.grad([r["lik"], r["y_hat"]], [yt, yb, m, s], [g_lik, g_y],
                                            allow_unused=True)
                out[f"{tag}/g_lik"] = g_lik.numpy()
                out[f"{tag}/g_yhat"] = g_y.numpy()
                for nm, gr in zip(("g_ytop", "g_ybase", "g_mu", "g_std"), grads):
                    out[f"{tag}/{nm}"] = (gr if gr is not None else torch.zeros(shape)).numpy()
    return out


def gen_gaussian(ref, gc):
    rng = np.random.default_rng(3003)
    out = {}
    shape = (2, 32, 4, 6)
    inputs = rng.normal(0, 3, size=shape).astype(np.float32)
    means = rng.normal(0, 1, size=shape).astype(np.float32)
    scales = np.exp(rng.normal(-1, 1.5, size=shape)).astype(np.float32)
    scales.ravel()[:6] = [0.0, 0.05, 0.11, 0.1100001, 1.0, 300.0]
    scales.ravel()[6:9] = [-1.0, 226.4, 255.9]
    inputs.ravel()[:4] = [0.5, 1.5, -2.5, 40.0]
    out["inputs"], out["means"], out["scales"] = inputs, means, scales
    out["scale_table"] = gc.scale_table.numpy()
    for use_means in (False, True):
        for training in (False, True):
            tag = f"{'means' if use_means else 'nomeans'}/{'train' if training else 'eval'}"
            x, s = t(inputs).requires_grad_(True), t(scales).requires_grad_(True)
            mu = t(means).requires_grad_(True) if use_means else None
            torch.manual_seed(5)
            o, lik = gc(x, s, mu, training=training)
            if training:
                torch.manual_seed(5)
                out[f"{tag}/noise"] = torch.empty(shape).uniform_(-0.5, 0.5).numpy()
            out[f"{tag}/outputs"], out[f"{tag}/lik"] = o.detach().numpy(), lik.detach().numpy()
            g = np.random.default_rng(9)
            g_o, g_l = (t(g.normal(0, 1, size=shape).astype(np.float32)) for _ in range(2))
            wrt = [x, s] + ([mu] if use_means else [])
            grads = torch.autograd.grad([o, lik], wrt, [g_o, g_l], allow_unused=True)
            out[f"{tag}/g_out"], out[f"{tag}/g_lik"] = g_o.numpy(), g_l.numpy()
            for nm, gr in zip(("g_inputs", "g_scales", "g_means"), grads):
                out[f"{tag}/{nm}"] = (gr if gr is not None else torch.zeros(shape)).numpy()
    out["likelihood/nomeans"] = gc._likelihood(t(inputs), t(scales)).numpy()
    out["likelihood/means"] = gc._likelihood(t(inputs), t(scales), t(means)).numpy()
    out["build_indexes"] = gc.build_indexes(t(scales)).numpy().astype(np.int32)
    probe = np.asarray([0, .05, .11, .1100001, 1.0, 226.4, 300., np.nan, np.inf, -np.inf], np.float32)
    out["build_indexes_probe/in"] = probe
    out["build_indexes_probe/out"] = gc.build_indexes(t(probe)).numpy().astype(np.int32)
    out["quantize/dequantize/nomeans"] = gc.quantize(t(inputs), "dequantize").numpy()
    out["quantize/dequantize/means"] = gc.quantize(t(inputs), "dequantize", t(means)).numpy()
    out["quantize/symbols/nomeans"] = gc.quantize(t(inputs), "symbols").numpy()
    out["quantize/symbols/means"] = gc.quantize(t(inputs), "symbols", t(means)).numpy()
    torch.manual_seed(6)
    out["quantize/noise/nomask"] = gc.quantize(t(inputs), "noise").numpy()
    mask = (rng.random(size=shape) < 0.5).astype(np.float32)
    out["quantize/mask"] = mask
    torch.manual_seed(6)
    out["quantize/noise/mask"] = gc.quantize(t(inputs), "noise", None, t(mask)).numpy()
    torch.manual_seed(6)
    out["quantize/noise/noise"] = torch.empty(shape).uniform_(-0.5, 0.5).numpy()
    sym = gc.quantize(t(inputs), "symbols", t(means))
    out["dequantize/means"] = ref.GaussianConditional.dequantize(sym, t(means)).numpy()
    out["dequantize/nomeans"] = ref.GaussianConditional.dequantize(sym).numpy()
    # known answers
    z = torch.zeros(1, 32, 2, 2)
    out["kat/masked_lik"] = gc(z, z, training=False)[1].numpy().ravel()[:1]
    out["kat/round"] = torch.round(torch.tensor([.5, 1.5, 2.5, -.5, -1.5])).numpy()
    return out


def main():
    ref = ref_shim.load_reference()
    gc = ref_shim.make_gaussian_conditional(ref)
    torch.set_num_threads(1)
    os.makedirs(OUT, exist_ok=True)
    files = {
        "masks.npz": gen_masks(ref),
        "large_quantiles.npz": gen_large_quantiles(),
        "slices.npz": gen_slices(ref, gc),
        "gaussian.npz": gen_gaussian(ref, gc),
    }
    manifest = {"torch": torch.__version__, "numpy": np.__version__,
                "generator": "oracle/gen_golden.py", "reference": "das-ankur/Efficient-PIC-with-Variance-Aware-Masking",
                "files": {}}
    for fname, arrays in files.items():
        path = os.path.join(OUT, fname)
        np.savez_compressed(path, **arrays)
        manifest["files"][fname] = {"arrays": len(arrays), "bytes": os.path.getsize(path)}
        print(fname, len(arrays), os.path.getsize(path))
    np.save(os.path.join(OUT, "scale_table.npy"), gc.scale_table.numpy())
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1)


if __name__ == "__main__":
    main()
