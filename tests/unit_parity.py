"""Shared harness: one SpatialTemporalConv (ours) against the fp64 oracle at an arbitrary shape, with the north-star
tolerance (1e-4, no noise-scaled slack) on the output, dx and every parameter gradient.

ReLU ties.  relu(z) = z * [z > 0]; the network is piecewise linear in its ReLU brackets.  At the BASELINE layer shapes a
unit has ~10^7 ReLU inputs, so a handful of them always sit within an fp32 rounding error of zero; there the bracket of ANY
fp32 implementation (ours, or the reference's own cuDNN path) differs from the fp64 one, and one flipped bracket moves
whole gradient tensors by O(1e-2) in the max norm although every kernel is right to 1e-7.  The harness therefore
  1. checks the forward output against the plain fp64 oracle (the forward is continuous across a tie),
  2. checks that our brackets differ from the fp64 brackets ONLY where the fp64 pre-activation is within `tie` (relative
     to its largest magnitude) of zero -- a genuinely wrong mask would fail here,
  3. compares every gradient with the fp64 oracle evaluated on OUR brackets (oracle `masks=`), i.e. on the same linear
     piece, at the full 1e-4 tolerance.
Reference: torch_src/models/mmargcn/agcn.py:96-136.  The oracle runs on the same device as the module (fp64 on the GPU
for the BASELINE shapes, CPU for the host-logic test)."""
import re

import numpy as np
import torch

from helpers import ZERO_GRAD, rel_err
from oracle import agcn_oracle as O

TOL = 1e-4
TIE = 1e-5


def loud_init(unit, seed):
    """BN gamma ~ U(.5, 1.5), beta ~ U(-.2, .2), adj_b ~ N(0, .1), conv biases ~ N(0, .05)."""
    g = torch.Generator().manual_seed(seed)
    for name, prm in unit.named_parameters():
        bn_like = re.search(r"(\.bn|down\.1)\.(weight|bias)$", name)
        if bn_like and name.endswith("weight"):
            prm.data = torch.rand(prm.shape, generator=g) + 0.5
        elif bn_like:
            prm.data = torch.rand(prm.shape, generator=g) * 0.4 - 0.2
        elif name.endswith("adj_b") or name.endswith("PA"):
            prm.data = torch.randn(prm.shape, generator=g) * 0.1
        elif name.endswith("bias"):
            prm.data = torch.randn(prm.shape, generator=g) * 0.05


def run_unit_parity(unit, x, w, device, tol=TOL, tie=TIE, b_name="adj_b", adj_a=None):
    """unit: a fresh SpatialTemporalConv / TCN_GCN_unit on CPU (already initialised); x: (N', C, T, V) fp32, w: upstream
    gradient.  Returns a dict of the measured errors; asserts the contract described in the module docstring."""
    stride, residual = unit.stride, unit._residual_kind
    state = {k: v.detach().clone() for k, v in unit.state_dict().items()}
    unit.to(device).train()
    xd = x.to(device).requires_grad_(True)
    wd = w.to(device)
    y = unit(xd)
    (y * wd).sum().backward()
    after = {k: v.detach().clone() for k, v in unit.state_dict().items()}       # running statistics after ONE training step
    with torch.no_grad():                      # our gcn-ReLU bracket (deterministic kernels: same o as inside the unit)
        o_ours = unit.gcn1(xd.detach())
    mask_o, mask_out = o_ours > 0, y.detach() > 0

    def leaves():
        p = O.as_leaves({"u." + k: v.to(device) for k, v in state.items()}, torch.float64)
        return p
    a64 = None if adj_a is None else torch.as_tensor(adj_a, dtype=torch.float64, device=device)
    # 1. forward against the plain fp64 oracle
    p_true, pre = leaves(), {}
    with torch.no_grad():
        y64, attn64 = O.st_unit(xd.detach().double(), p_true, "u", stride, residual, True, adj_a=a64, b_name=b_name, collect=pre)
    err = {"y": rel_err(y, y64)}
    assert err["y"] <= tol, f"output error {err['y']:.3e}"
    for k in range(3):
        e = rel_err(unit.gcn1.adj_c[k], attn64[k])
        assert e <= tol, f"adj_c[{k}] error {e:.3e}"
    # 2. brackets differ only at ties
    flips = 0
    for ours, key in ((mask_o, "pre_o"), (mask_out, "pre_out")):
        z = pre[key]
        diff = ours != (z > 0)
        flips += int(diff.sum())
        if diff.any():
            worst = float(z[diff].abs().max() / z.abs().max())
            assert worst <= tie, f"{key}: ReLU bracket differs where the fp64 pre-activation is {worst:.2e} (relative) from zero"
    err["relu_ties"] = flips
    # 3. gradients on our linear piece
    p64 = leaves()
    x64 = xd.detach().double().requires_grad_(True)
    y64m, _ = O.st_unit(x64, p64, "u", stride, residual, True, adj_a=a64, b_name=b_name, masks=(mask_o, mask_out))
    (y64m * wd.double()).sum().backward()
    err["dx"] = rel_err(xd.grad, x64.grad)
    assert err["dx"] <= tol, f"dx error {err['dx']:.3e}"
    ref = {k[2:]: v.grad for k, v in p64.items() if v.requires_grad}
    scale = max(float(v.abs().max()) for v in ref.values())
    worst = ("", 0.0)
    for name, prm in unit.named_parameters():
        r = ref[name]
        assert prm.grad is not None, name
        if ZERO_GRAD.search(name):
            e = float(prm.grad.abs().max()) / scale                     # mathematically zero: absolute bound
        else:
            denom = max(float(r.abs().max()), 1e-7 * scale)             # a gradient 1e7 below the unit's largest: absolute floor
            e = float((prm.grad.double() - r).abs().max()) / denom
        assert e <= tol, f"gradient {name} error {e:.3e}"
        if e > worst[1]:
            worst = (name, e)
    err["worst_grad"] = worst
    # running statistics after one training step
    new = after
    for k, v in p_true.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            a, b = new[k[2:]].double(), v
            e = float((a - b).abs().max()) / max(float(b.abs().max()), 1e-3)
            assert e <= tol, f"{k[2:]} error {e:.3e}"
    return err


def baseline_unit(M, G, cin, cout, stride, residual, v, seed):
    if v == 25:
        graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
    elif v == 22:
        graph = G.imu_fusion_graph(G.SkeletonGraph(G.MMACT_EDGES, center_joint=G.MMACT_CENTER), 4, "append_center", interconnect=True)
    elif v == 20:
        graph = G.SkeletonGraph(G.UTD_EDGES, center_joint=G.UTD_CENTER)
    else:
        raise ValueError(v)
    torch.manual_seed(seed)
    unit = M.SpatialTemporalConv(cin, cout, G.adjacency_from_graph(graph), stride=stride, residual=residual)
    loud_init(unit, seed + 1)
    return unit


def unit_inputs(nb, cin, cout, t, v, stride, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(nb, cin, t, v, generator=g)
    w = torch.randn(nb, cout, (t - 1) // stride + 1, v, generator=g)
    return x, w


def run_model_parity(model, state, x, w, device, num_channels, start, tol=TOL, tie=TOL, variant="mmargcn", adj_a=None):
    """Whole Model (ours) against the fp64 oracle with the same three-step contract as run_unit_parity: logits against the plain
    oracle, every unit's two ReLU brackets equal to the fp64 ones except at ties, all parameter gradients against the oracle
    evaluated on our brackets -- 1e-4, no noise-scaled slack.  ``model`` must be freshly loaded from ``state``.
    The tie window equals the forward tolerance here: through ten units the permitted forward error (1e-4 of the layer
    maximum) is exactly what may move a pre-activation across zero, so a bracket difference inside that window is explained by
    the forward contract and one outside it is a wrong mask."""
    from torch import nn
    units = [m for m in model.layers if not isinstance(m, nn.Dropout)]
    seen = []
    # Model(dropout > 0): the masks torch draws in the nn.Dropout(inplace=True) modules between our units are captured with hooks
    # and replayed in the oracle (reference layout), so logits and gradients are compared on identical masks
    drops, drop_in, drop_masks = [m for m in model.layers if isinstance(m, nn.Dropout)], [], []
    for d in drops:
        d.register_forward_pre_hook(lambda mod, inp: drop_in.append(inp[0].detach().clone()))
        d.register_forward_hook(lambda mod, inp, out: drop_masks.append(
            torch.where(drop_in[-1] != 0, out.detach() / drop_in[-1], torch.zeros_like(out)).permute(0, 3, 1, 2).double()))

    def wrap(unit):
        inner = unit.forward_cl

        def recording(h, **kw):
            out = inner(h)                                                # (the fused pooled tail is bypassed: the harness needs the feature map)
            seen.append((unit, h.detach(), out.detach().clone()))          # (a following in-place Dropout overwrites `out`)
            if kw.get("pool_groups"):
                from fusion_gcn_b200 import functional as FN
                return FN.PoolFn.apply(out, kw["pool_groups"])
            return out
        unit.forward_cl = recording
    for u in units:
        wrap(u)
    model.to(device).train()
    xd, wd = x.to(device), w.to(device)
    y = model(xd)
    (y * wd).sum().backward()
    after = {k: v.detach().clone() for k, v in model.state_dict().items()}
    assert len(drop_masks) == len(drops)
    dm = dict(dropout_masks=drop_masks) if drops else {}
    masks = []
    with torch.no_grad():
        for unit, h_in, out in seen:
            o = unit.gcn1.forward_cl(h_in)                               # (N', T, V, C) -> reference layout (N', C, T, V)
            masks.append(((o > 0).permute(0, 3, 1, 2), (out > 0).permute(0, 3, 1, 2)))
    a64 = None if adj_a is None else torch.as_tensor(adj_a, dtype=torch.float64, device=device)

    def leaves():
        return O.as_leaves({k: v.to(device) for k, v in state.items()}, torch.float64)
    p_true, pre = leaves(), []
    with torch.no_grad():
        y64 = O.model_forward(xd.double(), p_true, num_channels, True, start=start, variant=variant, adj_a=a64, collect=pre, **dm)
    err = {"y": rel_err(y, y64)}
    assert err["y"] <= tol, f"logit error {err['y']:.3e}"
    flips = 0
    for i, ((mo, mout), pr) in enumerate(zip(masks, pre)):
        for ours, key in ((mo, "pre_o"), (mout, "pre_out")):
            z = pr[key]
            diff = ours != (z > 0)
            flips += int(diff.sum())
            if diff.any():
                worst = float(z[diff].abs().max() / z.abs().max())
                assert worst <= tie, f"unit {i} {key}: ReLU bracket differs where the fp64 pre-activation is {worst:.2e} (relative) from zero"
    err["relu_ties"] = flips
    p64 = leaves()
    y64m = O.model_forward(xd.double(), p64, num_channels, True, start=start, variant=variant, adj_a=a64, masks=masks, **dm)
    (y64m * wd.double()).sum().backward()
    ref = {k: v.grad for k, v in p64.items() if v.requires_grad}
    scale = max(float(v.abs().max()) for v in ref.values())
    worst = ("", 0.0)
    for name, prm in model.named_parameters():
        r = ref[name]
        if ZERO_GRAD.search(name):
            e = float(prm.grad.abs().max()) / scale
        else:
            e = float((prm.grad.double() - r).abs().max()) / max(float(r.abs().max()), 1e-7 * scale)
        assert e <= tol, f"gradient {name} error {e:.3e}"
        if e > worst[1]:
            worst = (name, e)
    err["worst_grad"] = worst
    for k, v in p_true.items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            e = float((after[k].double() - v).abs().max()) / max(float(v.abs().max()), 1e-3)
            assert e <= tol, f"{k} error {e:.3e}"
    return err
