"""CPU: the C oracle (oracle/pic_oracle.c) against the golden vectors produced by the
reference's own Python functions (oracle/gen_golden.py).  This is what pins the oracle."""
import numpy as np
import pytest

import pic_oracle as po
from _common import (PR_LIST, assert_grad_close, assert_lik_close, golden, hashed_std, scale_table,
                     unpack_mask)


@pytest.fixture(scope="module", autouse=True)
def _built():
    po.build()


MASK_CASES = ["trained_n512", "trained_n1120", "trained_n8192", "modellike_n2048", "ties_n2048",
              "allequal_n512", "zeros_denormals_n512", "nan_n512", "inf_n512"]


@pytest.mark.parametrize("case", MASK_CASES)
def test_channel_mask_and_threshold(case):
    G = golden("masks.npz")
    std = G[f"{case}/std"]
    B = std.shape[0]
    flat = std.reshape(B, -1)
    for pr in PR_LIST:
        mask, thr = po.channel_mask(flat, pr)
        ref = unpack_mask(G[f"{case}/mask/pr={pr!r}"], flat.shape)
        assert np.array_equal(mask, ref), (case, pr)
        if 0 < pr < 10:
            ref_thr = G[f"{case}/thr/pr={pr!r}"]
            assert np.array_equal(thr, ref_thr, equal_nan=True), (case, pr, thr, ref_thr)


def test_progmask():
    G = golden("masks.npz")
    blocks = G["progmask/std"]  # [10,1,32,h,w]
    flat = blocks.reshape(blocks.shape[0], -1)
    for pr in PR_LIST:
        mask, _ = po.channel_mask(flat, pr)
        shape = tuple(G[f"progmask/shape/pr={pr!r}"])
        assert shape == (10, 32, 4, 6)
        ref = unpack_mask(G[f"progmask/mask/pr={pr!r}"], flat.shape)
        assert np.array_equal(mask, ref), pr


def test_large_quantiles_small_member():
    """n=49152 member of the regenerable large set (the 8M/16M members run in the gpu suite
    and in test_large_quantiles_slow)."""
    G = golden("large_quantiles.npz")
    x = hashed_std(49152, 5)
    for pr in (0.5, 1, 2.5, 5, 9.9999, 1e-4):
        thr, a, b = po.quantile(x, np.float32(1.0 - pr * 0.1))
        ref = G[f"n=49152/seed=5/pr={pr!r}"]
        assert (thr, a, b) == (ref[0], ref[1], ref[2]), pr
        assert int((x >= thr).sum()) == int(ref[3]) + 65536 * int(ref[4])


def test_large_quantile_f32_rank():
    """n = 8388608: rank = f32(q)*f32(n-1) differs from the f64 product."""
    G = golden("large_quantiles.npz")
    x = hashed_std(8388608, 7)
    for pr in (1, 9.9999):
        thr, a, b = po.quantile(x, np.float32(1.0 - pr * 0.1))
        ref = G[f"n=8388608/seed=7/pr={pr!r}"]
        assert (thr, a, b) == (ref[0], ref[1], ref[2]), pr


def test_quantile_too_large():
    with pytest.raises(RuntimeError, match="too large"):
        po.quantile(np.zeros((1 << 24) + 1, dtype=np.float32), 0.5)


SLICE_CASES = ["trained_n512", "trained_n2048", "model_n2048", "trained_n3072"]


@pytest.mark.parametrize("case", SLICE_CASES)
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_slice_forward_backward(case, mode):
    G = golden("slices.npz")
    table = scale_table()
    y_top, y_base, mu, std = (G[f"{case}/{k}"] for k in ("y_top", "y_base", "mu", "std"))
    B = std.shape[0]
    f = lambda a: a.reshape(B, -1)  # noqa: E731
    for pr in (0, 0.5, 1, 5, 7.3, 10):
        tag = f"{case}/{mode}/pr={pr!r}"
        noise = f(G[f"{case}/noise/pr={pr!r}"]) if mode == "train" else None
        o = po.slice_forward(f(y_top), f(y_base), f(mu), f(std), pr, table, noise=noise)
        mask = unpack_mask(G[f"{tag}/mask"], f(std).shape)
        assert np.array_equal(o["mask"], mask), tag
        assert np.array_equal(o["idx"], f(G[f"{tag}/idx"])), tag
        assert np.array_equal(o["symbols"], f(G[f"{tag}/symbols"])), tag
        assert np.array_equal(o["y_hat"], f(G[f"{tag}/y_hat"])), tag
        assert_lik_close(o["lik"], f(G[f"{tag}/lik"]), tag)
        ref_sum = float(G[f"{tag}/logsum"])
        assert abs(o["rate"].sum() - ref_sum) <= 1e-5 * abs(ref_sum) + 1e-6, tag
        g = po.slice_backward(f(G[f"{tag}/g_lik"]), f(G[f"{tag}/g_yhat"]), f(y_top), f(y_base), f(mu),
                              f(std), mask, noise)
        for k, gk in (("g_ytop", "g_ytop"), ("g_ybase", "g_ybase"), ("g_mu", "g_mu"), ("g_scale", "g_std")):
            assert_grad_close(g[k], f(G[f"{tag}/{gk}"]), f"{tag}/{k}")


@pytest.mark.parametrize("use_means", [False, True])
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_gaussian_conditional(use_means, mode):
    G = golden("gaussian.npz")
    inputs, means, scales = G["inputs"], G["means"], G["scales"]
    tag = f"{'means' if use_means else 'nomeans'}/{mode}"
    noise = G[f"{tag}/noise"] if mode == "train" else None
    mu = means if use_means else None
    out, lik = po.gaussian_forward(inputs, scales, mu, noise)
    assert np.array_equal(out, G[f"{tag}/outputs"])
    assert_lik_close(lik, G[f"{tag}/lik"], tag)
    g = po.gaussian_backward(G[f"{tag}/g_out"], G[f"{tag}/g_lik"], inputs, scales, mu, noise)
    assert_grad_close(g["g_inputs"], G[f"{tag}/g_inputs"], tag + "/g_inputs")
    assert_grad_close(g["g_scales"], G[f"{tag}/g_scales"], tag + "/g_scales")
    if use_means:
        assert_grad_close(g["g_means"], G[f"{tag}/g_means"], tag + "/g_means")


def test_build_indexes_quantize_kats():
    G = golden("gaussian.npz")
    table = G["scale_table"]
    assert np.array_equal(table, scale_table())
    assert np.array_equal(po.build_indexes(G["scales"], table), G["build_indexes"])
    assert np.array_equal(po.build_indexes(G["build_indexes_probe/in"], table), G["build_indexes_probe/out"])
    inputs, means = G["inputs"], G["means"]
    assert np.array_equal(po.quantize(inputs, "dequantize"), G["quantize/dequantize/nomeans"])
    assert np.array_equal(po.quantize(inputs, "dequantize", means), G["quantize/dequantize/means"])
    assert np.array_equal(po.quantize(inputs, "symbols"), G["quantize/symbols/nomeans"])
    assert np.array_equal(po.quantize(inputs, "symbols", means), G["quantize/symbols/means"])
    nz = G["quantize/noise/noise"]
    assert np.array_equal(po.quantize(inputs, "noise", noise=nz), G["quantize/noise/nomask"])
    assert np.array_equal(po.quantize(inputs, "noise", noise=nz, mask=G["quantize/mask"]), G["quantize/noise/mask"])
    with pytest.raises(ValueError):
        po.quantize(inputs, "bogus")
    # known answers recorded in SURVEY 8(c)
    z = np.zeros((1, 8), np.float32)
    assert po.gaussian_forward(z, z)[1][0, 0] == G["kat/masked_lik"][0] == np.float32(0.9999945163726807)
    assert np.array_equal(G["kat/round"], np.asarray([0, 2, 2, -0.0, -2], np.float32))


@pytest.mark.parametrize("q_name,pr", [("pr2.5", 2.5), ("pr5", 5), ("pr10", 10)])
def test_model_derived_config1(q_name, pr):
    """BASELINE config[0]: tensors captured from the random-init reference model on a 256x256 image
    (oracle/gen_golden_model.py): the oracle reproduces what the reference's per-slice code produced."""
    G = golden("model_c1.npz")
    table = scale_table()
    for k in range(10):
        f = lambda a: a.reshape(1, -1)  # noqa: E731
        y_top, y_base = G[f"pr2.5/slice{k}/y_top"], G[f"pr2.5/slice{k}/y_base"]
        mu, std = G[f"{q_name}/slice{k}/mu"], G[f"{q_name}/slice{k}/std"]
        assert std.shape == (1, 32, 16, 16)
        o = po.slice_forward(f(y_top), f(y_base), f(mu), f(std), pr, table)
        assert np.array_equal(o["mask"], unpack_mask(G[f"{q_name}/slice{k}/mask"], (1, std.size))), (q_name, k)
        assert np.array_equal(o["y_hat"], f(G[f"{q_name}/slice{k}/y_hat"])), (q_name, k)
        assert_lik_close(o["lik"], f(G[f"{q_name}/slice{k}/lik"]), f"{q_name}/slice{k}")


# ------------------------------------------------------------------------------------------ REM model (config[3])
def _rem_records(G, q_name):
    masks = [(G[f"{q_name}/mask{i}/scale"], float(G[f"{q_name}/mask{i}/pr"]), G[f"{q_name}/mask{i}/mask"])
             for i in range(int(G[f"{q_name}/n_mask"]))]
    return masks


@pytest.mark.parametrize("q_name", ["pr2.5", "pr5"])
def test_rem_model_masking_calls(q_name):
    """Every masking() call the random-init reference REM model made (checkpoint pass at q = 0.75, bar / star masks of
    apply_latent_enhancement, block mask on the REM-refined scale; models/rem_pic.py:182-189, 382-391, 583-586):
    the oracle reproduces each mask bit for bit."""
    G = golden("model_rem.npz")
    masks = _rem_records(G, q_name)
    assert len(masks) == 40 and int(G[f"{q_name}/n_mask_checkpoint"]) == 10
    prs = set()
    for scale, pr, packed in masks:
        flat = scale.reshape(scale.shape[0], -1)
        got, _ = po.channel_mask(flat, pr)
        assert np.array_equal(got, unpack_mask(packed, flat.shape)), pr
        prs.add(round(pr, 6))
    assert 0.75 in prs and float(G[f"{q_name}/pr"]) in prs      # checkpoint / bar quality and the requested one


def test_rem_model_entropy_calls():
    """gaussian_conditional(...) and build_indexes(...) as the REM model called them."""
    G = golden("model_rem.npz")
    table = scale_table()
    n_gc = int(G["pr2.5/n_gc"])
    assert n_gc == 20
    for i in range(n_gc):
        inp, sc = G[f"pr2.5/gc{i}/inputs"], G[f"pr2.5/gc{i}/scales"]
        means = G[f"pr2.5/gc{i}/means"] if f"pr2.5/gc{i}/means" in G.files else None
        f = lambda a: None if a is None else a.reshape(1, -1)  # noqa: E731
        out, lik = po.gaussian_forward(f(inp), f(sc), f(means))
        assert np.array_equal(out, f(G[f"pr2.5/gc{i}/outputs"])), i
        assert_lik_close(lik, f(G[f"pr2.5/gc{i}/lik"]), f"rem gc{i}")
        assert np.array_equal(po.build_indexes(f(sc), table), f(G[f"pr2.5/gc{i}/idx"])), i
    for i in range(int(G["pr2.5/n_idx"])):
        sc = G[f"pr2.5/idx{i}/scales"].reshape(1, -1)
        assert np.array_equal(po.build_indexes(sc, table), G[f"pr2.5/idx{i}/idx"].reshape(1, -1)), i


@pytest.mark.parametrize("q_name", ["pr2.5", "pr7"])
def test_codec_loops_of_the_reference_model(q_name):
    """models/pic.py:809-820 (compress) and 942-948 (decompress) of the random-init reference model, run with the oracle
    coder (oracle/gen_golden_model_codec.py; the generator asserted that encoder and decoder agree on scale, mask, index
    and symbols): per progressive slice the oracle reproduces the block mask, the index of scale * mask, the symbols of
    (y - mu) * mask and the decoder's dequantised slice."""
    G = golden("model_codec.npz")
    table = scale_table()
    pr = float(G[f"{q_name}/pr"])
    for k in range(10):
        tag = f"{q_name}/slice{k}"
        scale = G[f"{tag}/scale"]
        flat = scale.reshape(1, -1)
        mask, _ = po.channel_mask(flat, pr)
        assert np.array_equal(mask, unpack_mask(G[f"{tag}/mask"], flat.shape)), k
        assert abs(mask.mean() - pr / 10) < 2e-3
        assert np.array_equal(po.build_indexes(flat * mask, table), G[f"{tag}/idx"].reshape(1, -1).astype(np.int32)), k
        sym = po.quantize(G[f"{tag}/quant_in"].reshape(1, -1), "symbols")
        assert np.array_equal(sym, G[f"{tag}/symbols"].reshape(1, -1)), k
        y_hat = sym.astype(np.float32) + G[f"{tag}/mu"].reshape(1, -1)       # entropy_models.py:161-168: type_as(means) + means
        assert np.array_equal(y_hat, G[f"{tag}/y_hat"].reshape(1, -1)), k
