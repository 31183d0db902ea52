"""1-D graph convolutions (fusion_gcn_b200/graphconv.py; reference torch_src/models/mmargcn/graph_convolution.py:12-113).

not gpu: the drop-in modules over the torch stage backend, and the oracle restatements, against the LIVE reference classes
         (strict state-dict load, fp64) on an IMU graph built by the reference's own build_imu_graph_adjacency.
gpu:     the modules on the device against the fp64 oracle at IMU-graph sizes (up to 652 nodes = 163 steps x 4 signals), plus the
         large-V kernels (batched FFMA GEMM, column softmax) entry point by entry point against oracle/stages.py."""
import numpy as np
import pytest
import torch

from helpers import rel_err
from oracle import agcn_oracle as O, ref_loader, stages as S


def imu_adjacency(steps, signals, kind):
    """The reference's IMU graph (imu_feature_models.py:11-61) restated: nodes = (time step, signal); all signals of a step are
    connected both ways, every node points to the same signal one step later.  'agcn' -> the 3-subset partition, 'stgcn' ->
    the row-normalised undirected adjacency with self loops."""
    v = steps * signals
    edges = []
    for i in range(0, v, signals):
        for j in range(signals):
            for k in range(j + 1, signals):
                edges += [(i + j, i + k), (i + k, i + j)]
        if i >= signals:
            edges += [(i - signals + k, i + k) for k in range(signals)]
    if kind == "agcn":
        return O.partition_adjacency(edges, v)
    a = np.zeros((v, v))
    e = np.asarray(edges)
    a[e[:, 0], e[:, 1]] = 1.0
    a = np.maximum(a, a.T) + np.eye(v)              # undirected, with self loops (util/graph.py:116-124 called with add_self_loops)
    return torch.from_numpy(a / a.sum(axis=1, keepdims=True)).float()


def loud(module, seed):
    g = torch.Generator().manual_seed(seed)
    for name, prm in module.named_parameters():
        if name in ("bn.weight", "down.1.weight", "residual.1.weight"):
            prm.data = (torch.rand(prm.shape, generator=g) + 0.5).to(prm)
        elif name == "adj_b":
            prm.data = (torch.randn(prm.shape, generator=g) * 0.1).to(prm)
        elif name.endswith("bias"):
            prm.data = (torch.randn(prm.shape, generator=g) * 0.05).to(prm)


def check_against(ref_fn, module, x, w, tol):
    xo = x.clone().requires_grad_(True)
    y = module(xo)
    (y * w).sum().backward()
    p = O.as_leaves({"m." + k: v.detach().double() if v.is_floating_point() else v for k, v in module.state_dict().items()})
    for k in list(p):
        if k.endswith(".adj"):
            p[k] = p[k].detach()
    xr = x.double().clone().requires_grad_(True)
    yr = ref_fn(xr, p)
    (yr * w.double()).sum().backward()
    assert rel_err(y, yr) <= tol and rel_err(xo.grad, xr.grad) <= tol
    scale = max(float(v.grad.abs().max()) for v in p.values() if v.requires_grad and v.grad is not None)
    for k, prm in module.named_parameters():
        r = p["m." + k].grad
        bias_before_bn = k.endswith("bias") and (k.startswith("conv_d") or k.startswith("down.0") or k.startswith("residual.0") or k.startswith("conv_a"))
        if bias_before_bn:          # mathematically zero: absolute bound
            assert float(prm.grad.abs().max()) <= tol * scale, k
        else:
            assert float((prm.grad.double() - r).abs().max()) <= tol * max(float(r.abs().max()), 1e-7 * scale), k


# --------------------------------------------------------------------------------------------- not gpu
@pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")
def test_modules_and_oracle_against_the_live_reference(torch_stage_backend, monkeypatch):
    ref_loader.load()
    from models.mmargcn import graph_convolution as RG
    from models.mmargcn.imu_feature_models import build_imu_graph_adjacency
    import fusion_gcn_b200.modules as MM
    from fusion_gcn_b200 import graphconv as GC
    monkeypatch.setattr(MM, "_prep", lambda t: t.contiguous())              # keep fp64 through the stage backend
    adj3, adj1 = build_imu_graph_adjacency((10, 4), 4, "agcn"), build_imu_graph_adjacency((10, 4), 4, "stgcn", normalization="row")
    assert np.allclose(adj3, imu_adjacency(10, 4, "agcn")) and torch.allclose(adj1, imu_adjacency(10, 4, "stgcn"), atol=1e-6)
    g = torch.Generator().manual_seed(0)
    for cin, cout in ((1, 16), (16, 16), (16, 32)):
        ref = RG.AGCNGraphConvolution(cin, cout, adj3).double()
        loud(ref, 1)
        ours = GC.AGCNGraphConvolution(cin, cout, adj3).double()
        ours.load_state_dict(ref.state_dict(), strict=True)
        assert [k for k, _ in ours.named_parameters()] == [k for k, _ in ref.named_parameters()]
        x, w = torch.randn(3, cin, 40, generator=g, dtype=torch.float64), torch.randn(3, cout, 40, generator=g, dtype=torch.float64)
        xr = x.clone().requires_grad_(True)
        yr = ref(xr)
        (yr * w).sum().backward()
        xo = x.clone().requires_grad_(True)
        yo = ours(xo)
        (yo * w).sum().backward()
        assert rel_err(yo, yr) <= 1e-10 and rel_err(xo.grad, xr.grad) <= 1e-10
        for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
            assert float((a.grad - b.grad).abs().max()) <= 1e-9, k
        p = {"m." + k: v.clone() for k, v in ref.state_dict().items()}
        ref.train()
        assert rel_err(O.agcn_graph_conv_1d(x, p, "m", True), ref(x)) <= 1e-12                 # the oracle restatement itself
    for cin, cout, res, kind in ((1, 8, False, "none"), (8, 8, True, "identity"), (8, 16, True, "conv")):
        ref = RG.STGCNGraphConvolution(cin, cout, adj1, residual=res).double()
        loud(ref, 2)
        ours = GC.STGCNGraphConvolution(cin, cout, adj1, residual=res).double()
        ours.load_state_dict(ref.state_dict(), strict=True)
        x, w = torch.randn(3, cin, 40, generator=g, dtype=torch.float64), torch.randn(3, cout, 40, generator=g, dtype=torch.float64)
        xr = x.clone().requires_grad_(True)
        yr = ref(xr)
        (yr * w).sum().backward()
        xo = x.clone().requires_grad_(True)
        yo = ours(xo)
        (yo * w).sum().backward()
        assert rel_err(yo, yr) <= 1e-10 and rel_err(xo.grad, xr.grad) <= 1e-10
        for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
            assert float((a.grad - b.grad).abs().max()) <= 1e-9, k
        p = {"m." + k: v.clone() for k, v in ref.state_dict().items()}
        assert rel_err(O.stgcn_graph_conv_1d(x, p, "m", True, kind), ref(x)) <= 1e-12


# --------------------------------------------------------------------------------------------- gpu
@pytest.fixture(scope="module")
def K():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from fusion_gcn_b200 import ops
    return ops


def rnd(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed + sum(shape)))


@pytest.mark.gpu
@pytest.mark.parametrize("nb,t,v,c", [(2, 1, 40, 16), (3, 1, 652, 64), (2, 2, 100, 12), (1, 1, 33, 3)])
def test_large_graph_kernels_against_the_stage_oracle(K, nb, t, v, c):
    ci = max(1, c // 4)
    e = rnd(nb, t, v, 6 * ci)
    kw = dict(groups=3, offa=0, stridea=2 * ci, offb=ci, strideb=2 * ci, width=ci, nchunk=1)
    s = K.joint_gram(e.cuda(), e.cuda(), **kw)
    assert rel_err(s, S.joint_gram(e.double(), e.double(), **kw)) <= 2e-6
    adj_a, adj_b = rnd(3, v, v, seed=1) * 0.1, rnd(3, v, v, seed=2) * 0.1
    scale = 1.0 / (ci * t)
    p, g = K.attention_fwd(s, adj_a.cuda(), adj_b.cuda(), scale)
    p_ref, g_ref = S.attention_fwd(s.double().cpu(), adj_a.double(), adj_b.double(), scale)
    assert rel_err(p, p_ref) <= 2e-6 and rel_err(g, g_ref) <= 2e-6
    assert torch.allclose(p.sum(dim=-2), torch.ones(nb, 3, v, device="cuda"), atol=1e-5)
    x = rnd(nb, t, v, c, seed=3)
    z = K.joint_mix(x.cuda(), g, width=c, mode=K.MIX_AGG_FWD)
    assert rel_err(z, S.joint_mix(x.double(), g_ref, width=c, mode=S.MIX_AGG_FWD)) <= 2e-6
    dz = rnd(nb, t, v, 3 * c, seed=4)
    base = rnd(nb, t, v, c, seed=5)
    dx = K.joint_mix(dz.cuda(), g, width=c, mode=K.MIX_AGG_BWD, out=base.cuda().clone(), accumulate=True)
    assert rel_err(dx, S.joint_mix(dz.double(), g_ref, width=c, mode=S.MIX_AGG_BWD) + base.double()) <= 2e-6
    kw = dict(groups=3, offa=0, stridea=0, offb=0, strideb=c, width=c, nchunk=1)
    dg = K.joint_gram(x.cuda(), dz.cuda(), **kw)
    dg_ref = S.joint_gram(x.double(), dz.double(), **kw)
    assert rel_err(dg, dg_ref) <= 2e-6
    ds, dadj = K.attention_bwd(dg, p, scale)
    ds_ref, dadj_ref = S.attention_bwd(dg_ref, p_ref, scale)
    assert rel_err(ds, ds_ref) <= 5e-6 and rel_err(dadj, dadj_ref) <= 2e-6
    de = K.joint_mix(e.cuda(), ds, width=ci, mode=K.MIX_SCORE_BWD)
    assert rel_err(de, S.joint_mix(e.double(), ds_ref, width=ci, mode=S.MIX_SCORE_BWD)) <= 5e-6
    mat = rnd(v, v, seed=6) * 0.1
    flat = x.reshape(nb * t, v, c)
    for tr in (False, True):
        out = K.node_mix(flat.cuda(), mat.cuda(), transpose=tr)
        assert rel_err(out, S.node_mix(flat.double(), mat.double(), transpose=tr)) <= 2e-6


@pytest.mark.gpu
@pytest.mark.parametrize("steps,signals,cin,cout,n", [(163, 4, 1, 64, 4), (50, 6, 64, 64, 4), (40, 3, 64, 128, 8)])
def test_agcn_graph_convolution_1d_on_gpu(K, steps, signals, cin, cout, n):
    from fusion_gcn_b200 import graphconv as GC
    adj = imu_adjacency(steps, signals, "agcn")
    v = steps * signals
    torch.manual_seed(1)
    m = GC.AGCNGraphConvolution(cin, cout, adj)
    loud(m, 3)
    m.cuda().train()
    x, w = rnd(n, cin, v).cuda(), rnd(n, cout, v, seed=1).cuda()
    check_against(lambda xr, p: O.agcn_graph_conv_1d(xr, p, "m", True), m, x, w, 1e-4)
    assert m.adj_c[0].shape == (n, v, v)


@pytest.mark.gpu
@pytest.mark.parametrize("cin,cout,residual,kind", [(1, 64, False, "none"), (64, 64, True, "identity"), (64, 128, True, "conv")])
def test_stgcn_graph_convolution_1d_on_gpu(K, cin, cout, residual, kind):
    from fusion_gcn_b200 import graphconv as GC
    adj = imu_adjacency(50, 6, "stgcn")
    torch.manual_seed(2)
    m = GC.STGCNGraphConvolution(cin, cout, adj, residual=residual)
    loud(m, 4)
    m.cuda().train()
    x, w = rnd(4, cin, 300).cuda(), rnd(4, cout, 300, seed=1).cuda()
    check_against(lambda xr, p: O.stgcn_graph_conv_1d(xr, p, "m", True, kind), m, x, w, 1e-4)
