"""EntropyBottleneck for z: the numpy oracle against the golden vectors of the reference class (CPU);
the CUDA kernels -- forward, and the backward incl. every parameter gradient -- against both (GPU)."""
import numpy as np
import pytest
import torch

import eb_oracle
from _common import assert_grad_close, assert_lik_close, golden

PARAM_NAMES = [f"_matrix{i}" for i in range(5)] + [f"_bias{i}" for i in range(5)] + [f"_factor{i}" for i in range(4)] + ["quantiles"]


def _params(G):
    return {n: G[n] for n in PARAM_NAMES}


def test_oracle_matches_the_reference_class():
    G = golden("bottleneck.npz")
    P = _params(G)
    out, lik = eb_oracle.forward(P, G["z"])
    assert np.array_equal(out, G["eval/outputs"])
    assert_lik_close(lik, G["eval/lik"], "bottleneck eval")
    out_t, lik_t = eb_oracle.forward(P, G["z"], noise=G["train/noise"])
    np.testing.assert_allclose(out_t, G["train/outputs"], rtol=0, atol=1e-6)
    assert_lik_close(lik_t, G["train/lik"], "bottleneck train")
    assert_lik_close(eb_oracle.likelihood(P, G["likelihood/in"]), G["likelihood/out"], "bottleneck _likelihood")


def test_drop_in_has_the_reference_state_dict():
    import pic_b200

    G = golden("bottleneck.npz")
    eb = pic_b200.EntropyBottleneck(6)
    names = {n for n, _ in eb.named_parameters()}
    assert names == set(PARAM_NAMES)
    assert "target" in dict(eb.named_buffers()) and "_quantized_cdf" in dict(eb.named_buffers())
    with torch.no_grad():
        for n, p in eb.named_parameters():
            p.copy_(torch.from_numpy(G[n]))
    np.testing.assert_allclose(float(eb.loss()), float(G["loss"]), rtol=1e-6)
    assert eb.update() and eb._quantized_cdf.shape[0] == 6 and int(eb._cdf_length.min()) >= 3


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["eval", "train"])
def test_bottleneck_kernels(mode):
    import pic_b200

    G = golden("bottleneck.npz")
    dev = torch.device("cuda:0")
    eb = pic_b200.EntropyBottleneck(6).to(dev)
    with torch.no_grad():
        for n, p in eb.named_parameters():
            p.copy_(torch.from_numpy(G[n]).to(dev))
    z = torch.from_numpy(G["z"]).to(dev).requires_grad_(True)
    if mode == "eval":
        outputs, lik = eb(z, training=False)
    else:
        torch.manual_seed(123)
        outputs, lik = eb(z, training=True)       # same RNG call and layout as the reference: CPU and CUDA streams differ,
        noise = torch.from_numpy(G["train/noise"]).to(dev)   # so the golden noise is injected for the comparison
        from pic_b200.entropy_models import _BottleneckFn
        nz = (outputs - z).detach()
        assert float(nz.min()) >= -0.5 and float(nz.max()) <= 0.5
        outputs, lik = _BottleneckFn.apply(z, noise, eb._get_medians().reshape(-1).contiguous(), 1e-9, eb.filters,
                                           *eb._raw_parameters())
    if mode == "eval":
        assert np.array_equal(outputs.detach().cpu().numpy(), G["eval/outputs"])
    else:
        np.testing.assert_allclose(outputs.detach().cpu().numpy(), G["train/outputs"], rtol=0, atol=1e-6)
    assert_lik_close(lik.detach().cpu().numpy(), G[f"{mode}/lik"], f"bottleneck {mode}")
    w_lik, w_out = torch.from_numpy(G["w_lik"]).to(dev), torch.from_numpy(G["w_out"]).to(dev)
    loss = (torch.log(lik) * w_lik).sum() + (outputs * w_out).sum()
    loss.backward()
    assert_grad_close(z.grad.cpu().numpy(), G[f"{mode}/g_z"], f"g_z {mode}")
    for n, p in eb.named_parameters():
        want = G[f"{mode}/g_{n}"]
        got = np.zeros_like(want) if p.grad is None else p.grad.cpu().numpy()
        # a parameter gradient is a sum over all positions of the channel (f32 partial sums in another order than
        # autograd's, with cancellation): relative 1e-4 plus 2e-5 of the largest entry
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=2e-5 * max(float(np.abs(want).max()), 1e-30), err_msg=f"g_{n} {mode}")
    # the oracle on a larger seeded case, incl. the un-quantised _likelihood entry
    rng = np.random.default_rng(5)
    zz = (rng.normal(size=(4, 6, 9, 7)) * 5).astype(np.float32)
    ref_out, ref_lik = eb_oracle.forward(_params(G), zz)
    out2, lik2 = eb(torch.from_numpy(zz).to(dev), training=False)
    assert np.array_equal(out2.detach().cpu().numpy(), ref_out)
    assert_lik_close(lik2.detach().cpu().numpy(), ref_lik, "bottleneck vs oracle")
    vals = torch.from_numpy(G["likelihood/in"]).to(dev)
    assert_lik_close(eb._likelihood(vals).detach().cpu().numpy(), G["likelihood/out"], "bottleneck _likelihood")
    if mode == "eval" and isinstance(eb.entropy_coder, pic_b200.codec.RansCoder):
        # compress / decompress of z through the module API (reference 509-526): the decoded values are the quantised z
        eb.update()
        zq = torch.from_numpy(zz[:, :, :4, :4]).to(dev) * 0.5
        strings = eb.compress(zq)
        assert len(strings) == zq.shape[0]
        back = eb.decompress(strings, zq.shape[2:])
        want = eb(zq, training=False)[0]
        assert torch.equal(back, want)
