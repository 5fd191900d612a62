"""TEST INFRASTRUCTURE ONLY -- tests/golden/bottleneck.npz from the UNMODIFIED reference EntropyBottleneck
(entropy_models/entropy_models.py:296-528; loaded through oracle/ref_shim.py): parameters, inputs, eval and training
forward (with the noise the reference drew), _likelihood, loss(), and the reference's autograd gradients of z, every
parameter and the quantiles.      python oracle/gen_golden_bottleneck.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden", "bottleneck.npz")


def main():
    ref = ref_shim.load_reference()
    EB = ref.entropy_models_module.EntropyBottleneck
    torch.manual_seed(7)
    C = 6
    eb = EB(C)
    with torch.no_grad():   # move away from the symmetric initialisation so that every parameter matters
        for name, p in eb.named_parameters():
            if name.startswith("_factor"):
                p.uniform_(-0.8, 0.8)
            elif name.startswith("_matrix"):
                p.add_(torch.randn_like(p) * 0.3)
            elif name == "quantiles":
                p[:, 0, 1] = torch.randn(C) * 0.7
    out = {n: p.detach().numpy().copy() for n, p in eb.named_parameters()}
    out["target"] = eb.target.numpy()
    z = (torch.randn(3, C, 4, 5) * 3.0).requires_grad_(True)
    out["z"] = z.detach().numpy()
    g = torch.Generator().manual_seed(11)
    w_lik = torch.randn(z.shape, generator=g)
    w_out = torch.randn(z.shape, generator=g)
    out["w_lik"], out["w_out"] = w_lik.numpy(), w_out.numpy()
    for mode, training in (("eval", False), ("train", True)):
        eb.zero_grad()
        if z.grad is not None:
            z.grad = None
        torch.manual_seed(123)
        outputs, lik = eb(z, training=training)
        out[f"{mode}/outputs"], out[f"{mode}/lik"] = outputs.detach().numpy(), lik.detach().numpy()
        if training:
            out["train/noise"] = (outputs - z).detach().numpy()
        loss = (torch.log(lik) * w_lik).sum() + (outputs * w_out).sum()
        loss.backward()
        out[f"{mode}/g_z"] = z.grad.detach().numpy().copy()
        for n, p in eb.named_parameters():
            out[f"{mode}/g_{n}"] = (torch.zeros_like(p) if p.grad is None else p.grad).detach().numpy().copy()
    vals = torch.randn(C, 1, 17) * 4
    out["likelihood/in"] = vals.numpy()
    out["likelihood/out"] = eb._likelihood(vals).detach().numpy()
    out["loss"] = eb.loss().detach().numpy()
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "arrays", len(out))


if __name__ == "__main__":
    main()
