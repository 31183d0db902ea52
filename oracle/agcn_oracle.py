"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the AGCN / MMARGCN hot path.

A functional, state-dict-driven restatement of the reference's adaptive
graph-convolution unit in the reference's own tensor layout (N', C, T, V).  It
is the checker for the CUDA path (``tests/``, ``__graft_entry__.smoke()``) and
the ``cpu_baseline`` / ``--impl reference`` leg of ``bench.py`` (kind "port").
Nothing under ``fusion_gcn_b200/`` imports it.

Pinning: the reference has no tests, golden vectors or fixtures of its own
(SURVEY.md section 4), so this oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF: ``oracle/make_golden.py`` imports the unmodified reference from
/root/reference (build container only), runs it on seeded inputs and commits
inputs/weights/outputs/gradients under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against those fixtures and,
where /root/reference exists, against the live reference modules.

The arithmetic lives in PyTorch/ATen (reference pins torch==1.6.0,
requirements.txt:13-14; conv2d / matmul / softmax / batch_norm semantics are
unchanged in torch 2.11).  The same ATen ops are called here so the CPU timing
of this port is representative of the reference's CPU path.

Reference map (all under /root/reference):
  partition_adjacency   util/partition_strategy.py:42-46, util/graph.py:74-79,116-124
  temporal_conv         torch_src/models/mmargcn/agcn.py:37-51   (agcn/agcn.py:38-52)
  spatial_graph_conv    torch_src/models/mmargcn/agcn.py:96-115  (agcn/agcn.py:95-113)
  st_unit               torch_src/models/mmargcn/agcn.py:118-136 (agcn/agcn.py:116-133)
  model_forward         torch_src/models/mmargcn/agcn.py:183-200 (agcn/agcn.py:165-191)
  agcn_graph_conv_1d    torch_src/models/mmargcn/graph_convolution.py:96-113
  stgcn_graph_conv_1d   torch_src/models/mmargcn/graph_convolution.py:36-53
  init_state            torch_src/models/mmargcn/agcn.py:18-34,55-94,139-181
"""
import math
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5          # nn.BatchNorm2d default, never overridden (agcn.py:44,78,83,150)
BN_MOMENTUM = 0.1


# --------------------------------------------------------------------------- graph
def partition_adjacency(edges, num_vertices: Optional[int] = None) -> np.ndarray:
    """'spatial' partition: A[0] = I, A[1] = column-normalised reversed edges,
    A[2] = column-normalised edges (float64).  Edges (i, j) point from i towards
    the centre joint.  Zero-degree columns are written as explicit zeros (the
    reference leaves them uninitialised)."""
    e = np.unique(np.asarray(edges, dtype=np.int64), axis=0)
    v = int(e.max()) + 1 if num_vertices is None else int(num_vertices)
    out = np.zeros((3, v, v), dtype=np.float64)
    out[0] = np.eye(v)
    fwd = np.zeros((v, v), dtype=np.float64)
    fwd[e[:, 0], e[:, 1]] = 1.0
    for slot, adj in ((1, fwd.T.copy()), (2, fwd)):
        deg = adj.sum(axis=0)
        inv = np.zeros_like(deg)
        inv[deg > 0] = 1.0 / deg[deg > 0]
        out[slot] = adj * inv[None, :]          # adj . diag(inv)
    return out


def imu_fusion_edges(edges, num_vertices: int, center_joint: int, num_imu_joints: int,
                     interconnect: bool = False):
    """torch_src/models/mmargcn/fusion.py:65-89, mode 'append_center'."""
    new = [(num_vertices + i, center_joint) for i in range(num_imu_joints)]
    if interconnect:
        for i in range(num_imu_joints):
            for j in range(i + 1, num_imu_joints):
                new.append((num_vertices + i, num_vertices + j))
    return np.vstack((np.asarray(edges), np.asarray(new, dtype=np.int64)))


# --------------------------------------------------------------------------- layers
def _bn(x, p, prefix, training):
    rm, rv = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    if training and (prefix + ".num_batches_tracked") in p:
        p[prefix + ".num_batches_tracked"] += 1
    return F.batch_norm(x, rm, rv, p[prefix + ".weight"], p[prefix + ".bias"], training, BN_MOMENTUM, BN_EPS)


def temporal_conv(x, p: Dict[str, torch.Tensor], prefix: str, stride: int, training: bool):
    w = p[prefix + ".conv.weight"]
    pad = (w.shape[2] - 1) // 2
    y = F.conv2d(x, w, p[prefix + ".conv.bias"], stride=(stride, 1), padding=(pad, 0))
    return _bn(y, p, prefix + ".bn", training)


def _relu(pre, mask, collect, key):
    """relu(pre).  ``collect`` (a dict) receives the pre-activation; ``mask`` (a bool tensor) REPLACES the sign test:
    relu(z) = z * [z > 0], and a parity test may fix the bracket to the mask the implementation under test used, so that
    gradients are compared on the same linear piece of the network (elements whose pre-activation is a rounding error
    away from zero otherwise flip the bracket and move whole gradient tensors by O(1e-2) -- in the fp32 reference too)."""
    if collect is not None:
        collect[key] = pre
    return torch.relu(pre) if mask is None else pre * mask.to(pre.dtype)


def spatial_graph_conv(x, p, prefix: str, training: bool, adj_a: Optional[torch.Tensor] = None,
                       b_name: str = "adj_b", mask=None, collect=None) -> Tuple[torch.Tensor, List[torch.Tensor]]:
    n, c, t, v = x.shape
    a = p[prefix + ".adj_a"] if adj_a is None else adj_a
    fixed_plus_learned = a + p[prefix + "." + b_name]
    num_subsets = fixed_plus_learned.shape[0]
    flat = x.reshape(n, c * t, v)
    acc = None
    attn = []
    for k in range(num_subsets):
        theta = F.conv2d(x, p[f"{prefix}.conv_a.{k}.weight"], p[f"{prefix}.conv_a.{k}.bias"])
        phi = F.conv2d(x, p[f"{prefix}.conv_b.{k}.weight"], p[f"{prefix}.conv_b.{k}.bias"])
        ci = theta.shape[1]
        th = theta.permute(0, 3, 1, 2).reshape(n, v, ci * t)
        ph = phi.reshape(n, ci * t, v)
        pk = torch.softmax(torch.matmul(th, ph) / (ci * t), dim=-2)      # columns sum to one
        attn.append(pk)
        mixed = torch.matmul(flat, pk + fixed_plus_learned[k]).reshape(n, c, t, v)
        z = F.conv2d(mixed, p[f"{prefix}.conv_d.{k}.weight"], p[f"{prefix}.conv_d.{k}.bias"])
        acc = z if acc is None else acc + z
    y = _bn(acc, p, prefix + ".bn", training)
    if (prefix + ".down.0.weight") in p:
        d = F.conv2d(x, p[prefix + ".down.0.weight"], p[prefix + ".down.0.bias"])
        d = _bn(d, p, prefix + ".down.1", training)
    else:
        d = x
    return _relu(y + d, mask, collect, "pre_o"), attn


def agcn_graph_conv_1d(x, p, prefix: str, training: bool):
    """AGCNGraphConvolution.forward, torch_src/models/mmargcn/graph_convolution.py:96-113: x (N, C, V); Conv1d weights
    (out, in, 1); softmax over dim -2 of theta^T phi / Ci; BatchNorm1d; down = identity | Conv1d + BatchNorm1d."""
    n, c, v = x.shape
    adj = p[prefix + ".adj_a"] + p[prefix + ".adj_b"]
    y = None
    for k in range(adj.shape[0]):
        theta = F.conv1d(x, p[f"{prefix}.conv_a.{k}.weight"], p[f"{prefix}.conv_a.{k}.bias"]).permute(0, 2, 1)      # (N, V, Ci)
        phi = F.conv1d(x, p[f"{prefix}.conv_b.{k}.weight"], p[f"{prefix}.conv_b.{k}.bias"])                         # (N, Ci, V)
        att = torch.softmax(torch.matmul(theta, phi) / theta.shape[-1], dim=-2) + adj[k]
        z = F.conv1d(torch.matmul(x, att), p[f"{prefix}.conv_d.{k}.weight"], p[f"{prefix}.conv_d.{k}.bias"])
        y = z if y is None else y + z
    y = _bn(y, p, prefix + ".bn", training)
    if (prefix + ".down.0.weight") in p:
        d = _bn(F.conv1d(x, p[prefix + ".down.0.weight"], p[prefix + ".down.0.bias"]), p, prefix + ".down.1", training)
    else:
        d = x
    return torch.relu(y + d)


def stgcn_graph_conv_1d(x, p, prefix: str, training: bool, residual: str):
    """STGCNGraphConvolution.forward, graph_convolution.py:36-53 (dense adjacency, no dropout): relu(conv(x) adj^T + residual(x)),
    residual in {'none', 'identity', 'conv'}."""
    support = F.conv1d(x, p[prefix + ".conv.weight"], p.get(prefix + ".conv.bias"))
    out = torch.matmul(support, p[prefix + ".adj"].t())
    if residual == "identity":
        out = out + x
    elif residual == "conv":
        out = out + _bn(F.conv1d(x, p[prefix + ".residual.0.weight"], p[prefix + ".residual.0.bias"]), p, prefix + ".residual.1", training)
    return torch.relu(out)


def st_unit(x, p, prefix: str, stride: int, residual: str, training: bool,
            adj_a=None, b_name="adj_b", masks=None, collect=None):
    """residual in {'none', 'identity', 'conv'}.  ``masks`` = (mask of the gcn ReLU, mask of the output ReLU) and
    ``collect`` as in :func:`_relu`."""
    o, attn = spatial_graph_conv(x, p, prefix + ".gcn1", training, adj_a, b_name,
                                 mask=None if masks is None else masks[0], collect=collect)
    u = temporal_conv(o, p, prefix + ".tcn1", stride, training)
    if residual == "identity":
        u = u + x
    elif residual == "conv":
        u = u + temporal_conv(x, p, prefix + ".residual", stride, training)
    return _relu(u, None if masks is None else masks[1], collect, "pre_out"), attn


# --------------------------------------------------------------------------- model
def layer_plan(num_channels: int, start: int = 64, num_layers: int = 10):
    """(cin, cout, stride, residual) rows of mmargcn/agcn.py:152-164."""
    rows = [(num_channels, start, 1, "none"), (start, start, 1, "identity"), (start, start, 1, "identity"),
            (start, start, 1, "identity"), (start, 2 * start, 2, "conv"), (2 * start, 2 * start, 1, "identity"),
            (2 * start, 2 * start, 1, "identity"), (2 * start, 4 * start, 2, "conv"),
            (4 * start, 4 * start, 1, "identity"), (4 * start, 4 * start, 1, "identity")]
    return rows[:min(len(rows), num_layers)]


def model_forward(x, p, num_channels: int, training: bool, start: int = 64, num_layers: int = 10,
                  variant: str = "mmargcn", adj_a=None, return_attention: bool = False, dropout_masks=None,
                  masks=None, collect=None):
    """x: (N, M, T, V, C).  variant 'mmargcn' -> layers l0.., param adj_b, buffer adj_a;
    variant 'original' -> layers l1.., param PA, adjacency passed as ``adj_a``.
    ``dropout_masks``: the (already 1/(1-p)-scaled) masks of the nn.Dropout modules the reference inserts after every
    unit but the last when dropout > 0 (agcn.py:166-169), each (N*M, C, T', V); the state-dict keys then follow the
    reference's renumbering (unit i is l{2i}, the dropouts take the odd names).
    ``masks`` / ``collect``: per-unit ReLU brackets / pre-activations as in :func:`st_unit` (lists with one entry per unit)."""
    n, m, t, v, c = x.shape
    h = x.permute(0, 1, 3, 4, 2).reshape(n, m * v * c, t)
    h = _bn(h, p, "data_bn", training)
    h = h.reshape(n, m, v, c, t).permute(0, 1, 3, 4, 2).reshape(n * m, c, t, v)
    first = 0 if variant == "mmargcn" else 1
    b_name = "adj_b" if variant == "mmargcn" else "PA"
    attention = []
    step = 1 if dropout_masks is None else 2
    for i, (_, _, stride, residual) in enumerate(layer_plan(num_channels, start, num_layers)):
        pre = None
        if collect is not None:
            pre = {}
            collect.append(pre)
        h, attn = st_unit(h, p, f"l{step * i + first}", stride, residual, training, adj_a, b_name,
                          masks=None if masks is None else masks[i], collect=pre)
        if dropout_masks is not None and i < len(dropout_masks):
            h = h * dropout_masks[i].to(h.dtype)
        attention.append(attn)
    feat = h.reshape(n, m, h.shape[1], -1).mean(3).mean(1)
    if "fc.weight" in p:
        feat = F.linear(feat, p["fc.weight"], p["fc.bias"])
    return (feat, attention) if return_attention else feat


# --------------------------------------------------------------------------- init
def _conv_param(g, cout, cin, k, std):
    return torch.randn(cout, cin, k, 1, generator=g) * std, torch.zeros(cout)


def init_state(adj: np.ndarray, data_shape, num_classes: int, start: int = 64, num_layers: int = 10,
               seed: int = 0, loud: bool = False, variant: str = "mmargcn", dtype=torch.float32):
    """State dict with the reference's key names and initial distributions
    (kaiming-normal fan_out convs, zero biases, BN gamma 1 except gcn bn 1e-6,
    adj_b 1e-6, conv_d ~ N(0, sqrt(2/(Cout*Cin*3)))).  The random stream is this
    function's own; parity tests copy one state dict into both sides.
    ``loud=True`` applies SURVEY D7's init (BN gamma~U(.5,1.5), beta~U(-.2,.2),
    adj_b~N(0,.1)) so that the adaptive branch is visible at the 1e-4 level."""
    g = torch.Generator().manual_seed(seed)
    m, _, v, c = data_shape
    a32 = torch.from_numpy(adj.astype(np.float32))
    k_sub = a32.shape[0]
    p = {}

    def bn(prefix, ch, gamma=1.0):
        p[prefix + ".weight"] = torch.full((ch,), gamma)
        p[prefix + ".bias"] = torch.zeros(ch)
        p[prefix + ".running_mean"] = torch.zeros(ch)
        p[prefix + ".running_var"] = torch.ones(ch)
        p[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        if loud:
            p[prefix + ".weight"] = torch.rand(ch, generator=g) + 0.5
            p[prefix + ".bias"] = torch.rand(ch, generator=g) * 0.4 - 0.2

    def conv(prefix, cout, cin, k=1, std=None):
        std = math.sqrt(2.0 / (cout * k)) if std is None else std        # kaiming fan_out
        p[prefix + ".weight"], p[prefix + ".bias"] = _conv_param(g, cout, cin, k, std)

    bn("data_bn", m * v * c)
    first = 0 if variant == "mmargcn" else 1
    b_name = "adj_b" if variant == "mmargcn" else "PA"
    for i, (cin, cout, stride, residual) in enumerate(layer_plan(c, start, num_layers)):
        pre = f"l{i + first}"
        p[f"{pre}.gcn1.{b_name}"] = (torch.randn(a32.shape, generator=g) * 0.1) if loud else torch.full_like(a32, 1e-6)
        if variant == "mmargcn":
            p[f"{pre}.gcn1.adj_a"] = a32.clone()
        ci = cout // 4
        for k in range(k_sub):
            conv(f"{pre}.gcn1.conv_a.{k}", ci, cin)
            conv(f"{pre}.gcn1.conv_b.{k}", ci, cin)
            conv(f"{pre}.gcn1.conv_d.{k}", cout, cin, std=math.sqrt(2.0 / (cout * cin * 1 * k_sub)))
        if cin != cout:
            conv(f"{pre}.gcn1.down.0", cout, cin)
            bn(f"{pre}.gcn1.down.1", cout)
        bn(f"{pre}.gcn1.bn", cout, 1e-6)
        conv(f"{pre}.tcn1.conv", cout, cout, 9)
        bn(f"{pre}.tcn1.bn", cout)
        if residual == "conv":
            conv(f"{pre}.residual.conv", cout, cin, 1)
            bn(f"{pre}.residual.bn", cout)
    last = layer_plan(c, start, num_layers)[-1][1]
    p["fc.weight"] = torch.randn(num_classes, last, generator=g) * math.sqrt(2.0 / num_classes)
    bound = 1.0 / math.sqrt(last)
    p["fc.bias"] = (torch.rand(num_classes, generator=g) * 2 - 1) * bound
    if loud:   # non-zero conv biases so that every bias path is exercised
        for name in list(p):
            if name.endswith(".bias") and ".conv" in name or name.endswith("down.0.bias"):
                p[name] = torch.randn(p[name].shape, generator=g) * 0.05
    return {k_: (t_.to(dtype) if t_.is_floating_point() else t_) for k_, t_ in p.items()}


def as_leaves(state: Dict[str, torch.Tensor], dtype=None):
    """Clone a state dict into autograd leaves (floating tensors that are not BN
    running statistics or adj_a require grad)."""
    out = {}
    for k, t in state.items():
        t = t.detach().clone()
        if t.is_floating_point():
            if dtype is not None:
                t = t.to(dtype)
            if not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("adj_a")):
                t.requires_grad_(True)
        out[k] = t
    return out


def rel_err(a: torch.Tensor, b_ref: torch.Tensor) -> float:
    """maxabs(a - ref) / maxabs(ref)  (SURVEY D8 metric)."""
    a = a.detach().double().cpu()
    b = b_ref.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)
