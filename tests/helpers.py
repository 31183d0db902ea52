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


def eval_mode_gradient_case(M, G, device, tol):
    """Gradients under model.eval() (frozen-BatchNorm fine-tuning, saliency maps): every BatchNorm is an affine map of its running
    statistics, so the biases of the convolutions in front of them get real gradients; compared with the fp64 oracle's autograd."""
    from oracle import agcn_oracle as O
    g = load_golden("model_utd_s8")
    m, t, v, c, ncls, start = [int(a) for a in g["meta"]]
    state = {k: to_t(a) for k, a in sub(g, "state.").items()}
    state.update({k: to_t(a).to(state[k].dtype) for k, a in sub(g, "f64.after.").items()})      # non-trivial running statistics
    model = M.Model((m, t, v, c), ncls, G.SkeletonGraph(G.UTD_EDGES), start_feature_size=start).to(device)
    model.load_state_dict(state, strict=True)
    model.eval()
    x = to_t(g["x"]).to(device).requires_grad_(True)
    w = to_t(g["w"]).to(device)
    y = model(x)
    (y * w).sum().backward()
    before = {k: b.clone() for k, b in model.named_buffers()}
    p = O.as_leaves({k: a.double() if a.is_floating_point() else a for k, a in state.items()})
    x64 = to_t(g["x"]).double().requires_grad_(True)
    y64 = O.model_forward(x64, p, c, False, start=start)
    (y64 * to_t(g["w"]).double()).sum().backward()
    assert rel_err(y, y64) <= tol
    assert rel_err(x.grad, x64.grad) <= tol
    ref = {k: p[k].grad for k, _ in model.named_parameters()}
    scale = max(float(r.abs().max()) for r in ref.values())
    for k, q in model.named_parameters():
        assert q.grad is not None, k
        # no analytic zeros here: tensors whose gradient is tiny are compared on the scale of the largest gradient
        err = float((q.grad.detach().double().cpu() - ref[k]).abs().max()) / max(float(ref[k].abs().max()), 1e-3 * scale)
        assert err <= tol, f"{k}: {err:.3e}"
    assert float(ref["l3.tcn1.conv.bias"].abs().max()) > 0 and float(model.l3.tcn1.conv.bias.grad.abs().max()) > 0
    for k, b in model.named_buffers():                                                             # eval mode leaves the statistics alone
        assert torch.equal(b, before[k]), k
    model(to_t(g["x"]).to(device))                                                                 # no_grad-free eval forward still works
    with torch.no_grad():
        assert rel_err(model(to_t(g["x"]).to(device)), y64) <= tol


def recompute_case(M, G, device, cin=8, cout=16, stride=2, v=25, t=12, nb=3):
    """One unit run twice from the same state, once keeping and once recomputing theta / phi and the aggregated tensor: output, dx and
    every parameter gradient must be IDENTICAL (the recomputation runs the same deterministic kernels on the same inputs)."""
    import copy
    torch.manual_seed(5)
    unit = M.SpatialTemporalConv(cin, cout, G.partition_adjacency(G.NTU_EDGES if v == 25 else G.UTD_EDGES), stride=stride).to(device).train()
    twin = M.set_recompute(copy.deepcopy(unit), True)
    assert twin.gcn1._agcn_recompute and twin._agcn_recompute and not unit._agcn_recompute
    x = torch.randn(nb, cin, t, v, device=device)
    w = torch.randn(nb, cout, (t - 1) // stride + 1, v, device=device)
    outs = []
    for mod in (unit, twin):
        xi = x.clone().requires_grad_(True)
        y = mod(xi)
        (y * w).sum().backward()
        outs.append((y.detach(), xi.grad, {k: p.grad for k, p in mod.named_parameters()}))
    (y0, dx0, g0), (y1, dx1, g1) = outs
    assert torch.equal(y0, y1) and torch.equal(dx0, dx1)
    for k in g0:
        assert torch.equal(g0[k], g1[k]), k



def second_backward_case(unit, x):
    """The activation store goes through save_for_backward: a second backward through the same forward works under retain_graph=True
    (gradients accumulate: twice the first) and raises torch's own error without it."""
    import pytest
    x = x.detach().clone().requires_grad_(True)
    unit.zero_grad(set_to_none=True)
    out = unit(x)
    out.sum().backward(retain_graph=True)
    first = {k: p.grad.clone() for k, p in unit.named_parameters() if p.grad is not None}
    dx = x.grad.clone()
    out.sum().backward()
    assert rel_err(x.grad, 2 * dx) <= 1e-6
    for k, p in unit.named_parameters():
        if k in first and float(first[k].abs().max()) > 0:
            assert rel_err(p.grad, 2 * first[k]) <= 1e-6, k
    with pytest.raises(RuntimeError, match="second time|already been freed"):
        out.sum().backward()
