"""Shared test helpers: golden-fixture loading and the SURVEY D8 gradient metric."""
import os
import re

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# parameter-gradient families that are mathematically zero: biases followed by a
# training-mode BN, and the theta bias (constant along the softmax axis)
ZERO_GRAD = re.compile(r"(conv_d\.\d\.bias|down\.0\.bias|tcn1\.conv\.bias|residual\.conv\.bias|conv_a\.\d\.bias)$")

UNIT_FIXTURES = ["first_c3_16", "same_16_16", "down_16_32_s2", "same_32_32_v22", "wide_c9_16"]
MODEL_FIXTURES = ["utd_s8", "ntu_s8_m2", "utd_s8_default_init"]
RESIDUAL_KINDS = {0: "none", 1: "identity", 2: "conv"}


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def sub(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


def to_t(a, dtype=None, device="cpu"):
    t = torch.from_numpy(np.asarray(a))
    if dtype is not None and t.is_floating_point():
        t = t.to(dtype)
    return t.to(device)


def rel_err(a, ref):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(ref).detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def check_grads(ours: dict, ref: dict, tol: float, what="", ref32: dict = None, slack: float = 4.0):
    """ours/ref: name -> gradient (ref = fp64 ground truth).  Non-degenerate tensors: maxabs(diff)/maxabs(ref) <= tol.
    Zero-gradient families: maxabs(ours) <= tol * largest reference gradient magnitude.
    ``ref32`` (the reference evaluated in fp32) widens the per-tensor bound to slack * its own error against fp64:
    on ill-conditioned seeded models the fp32 reference itself is far from 1e-4 (measured up to 1.4e-2), and parity
    with the reference cannot be tighter than the reference's own rounding noise."""
    scale = max(float(np.abs(np.asarray(torch.as_tensor(v).detach().cpu())).max()) for v in ref.values())
    worst = ("", 0.0)
    for name, r in ref.items():
        bound = tol
        if ref32 is not None and not ZERO_GRAD.search(name):
            bound = max(tol, slack * rel_err(ref32[name], r))
        assert name in ours, f"{what}: missing gradient {name}"
        g = ours[name]
        assert g is not None, f"{what}: gradient {name} is None"
        assert tuple(g.shape) == tuple(r.shape), f"{what}: {name} shape {tuple(g.shape)} vs {tuple(r.shape)}"
        if ZERO_GRAD.search(name):
            e = float(torch.as_tensor(g).detach().abs().max()) / scale
        else:
            e = rel_err(g, r)
        if e > worst[1]:
            worst = (name, e)
        assert e <= bound, f"{what}: gradient {name} error {e:.3e} > {bound:.1e}"
    return worst


def stat_err(a, ref, floor=1e-3):
    """Error metric for BN running statistics: maxabs(diff) / max(maxabs(ref), floor).  A running mean that is
    mathematically zero (e.g. a conv fed by zero-mean data_bn output) is compared on the absolute floor."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(ref).detach().double().cpu()
    return (a - b).abs().max().item() / max(b.abs().max().item(), floor)
