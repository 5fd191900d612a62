"""TEST INFRASTRUCTURE ONLY -- numpy (f32) restatement of the reference EntropyBottleneck's likelihood
(entropy_models/entropy_models.py:403-436, 449-492); pinned to tests/golden/bottleneck.npz, which the reference's own
class produced (oracle/gen_golden_bottleneck.py).  Never imported by the package."""
import numpy as np

f32 = np.float32


def _softplus(x):
    return np.where(x > 20, x, np.log1p(np.exp(x.astype(np.float64)))).astype(f32)


def logits_cumulative(params: dict, v: np.ndarray) -> np.ndarray:
    """v: [C, 1, N] -> logits [C, 1, N] (reference 403-420)."""
    logits = v.astype(f32)
    for i in range(5):
        logits = np.matmul(_softplus(params[f"_matrix{i}"]), logits).astype(f32) + params[f"_bias{i}"]
        if i < 4:
            logits = (logits + np.tanh(params[f"_factor{i}"]) * np.tanh(logits)).astype(f32)
    return logits


def _sigmoid(x):
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(f32)


def likelihood(params: dict, v: np.ndarray) -> np.ndarray:
    """Reference 422-434 on values [C, 1, N]."""
    lower = logits_cumulative(params, v - f32(0.5))
    upper = logits_cumulative(params, v + f32(0.5))
    sign = -np.sign(lower + upper)
    return np.abs(_sigmoid(sign * upper) - _sigmoid(sign * lower)).astype(f32)


def forward(params: dict, z: np.ndarray, noise=None, lik_bound: float = 1e-9):
    """Reference 449-492: z [B, C, ...] -> (outputs, likelihood), eval (noise None) or training."""
    B, C = z.shape[:2]
    values = np.moveaxis(z, 1, 0).reshape(C, 1, -1)
    med = params["quantiles"][:, :, 1:2]
    if noise is None:
        outputs = (np.round(values - med) + med).astype(f32)          # np.round: half to even, as torch.round
    else:
        outputs = (values + np.moveaxis(noise, 1, 0).reshape(C, 1, -1)).astype(f32)
    lik = np.maximum(likelihood(params, outputs), f32(lik_bound)) if lik_bound > 0 else likelihood(params, outputs)
    back = lambda a: np.moveaxis(a.reshape((C, B) + z.shape[2:]), 0, 1)  # noqa: E731
    return back(outputs), back(lik)
