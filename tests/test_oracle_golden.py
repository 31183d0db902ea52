"""Pins oracle/agcn_oracle.py: (1) against the committed golden vectors that oracle/make_golden.py
produced by running the UNMODIFIED reference, (2) against the live reference where /root/reference exists."""
import numpy as np
import pytest
import torch

from helpers import MODEL_FIXTURES, RESIDUAL_KINDS, UNIT_FIXTURES, check_grads, load_golden, rel_err, sub, to_t
from oracle import agcn_oracle as O
from oracle import ref_loader


def test_adjacency_matches_reference_golden():
    g = load_golden("adjacency")
    for name in ("ntu", "utd", "mmact"):
        a = O.partition_adjacency(g[name + "_edges"])
        assert np.array_equal(a, g[name + "_adj"]), name
    assert tuple(g["ntu_adj"].sum(axis=(1, 2))) == (25.0, 24.0, 20.0)
    for inter in (False, True):
        e = O.imu_fusion_edges(g["mmact_edges"], 18, int(g["mmact_center"]), 4, interconnect=inter)
        key = "mmact_imu4_" + ("inter" if inter else "plain") + "_adj"
        assert np.array_equal(O.partition_adjacency(e), g[key])


@pytest.mark.parametrize("name", UNIT_FIXTURES)
@pytest.mark.parametrize("tag,dtype,tol", [("f32", torch.float32, 2e-6), ("f64", torch.float64, 1e-12)])
def test_unit_oracle_vs_golden(name, tag, dtype, tol):
    g = load_golden("unit_" + name)
    cin, cout, stride, res = [int(v) for v in g["meta"]]
    state = {"u." + k: to_t(v, dtype) for k, v in sub(g, "state.").items()}
    p = O.as_leaves(state)
    x = to_t(g["x"], dtype).requires_grad_(True)
    y, attn = O.st_unit(x, p, "u", stride, RESIDUAL_KINDS[res], True)
    (y * to_t(g["w"], dtype)).sum().backward()
    assert rel_err(y, g[tag + ".y"]) <= tol
    assert rel_err(x.grad, g[tag + ".dx"]) <= tol * 10
    for k in range(3):
        assert rel_err(attn[k], g[f"{tag}.adj_c.{k}"]) <= tol
    ours = {k[2:]: v.grad for k, v in p.items() if v.requires_grad}
    check_grads(ours, sub(g, tag + ".grad."), tol * 50, name)
    for k, v in sub(g, tag + ".after.").items():
        assert rel_err(p["u." + k], v) <= tol


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_model_oracle_vs_golden(name):
    g = load_golden("model_" + name)
    m, t, v, c, ncls, start = [int(a) for a in g["meta"]]
    dtype = torch.float64
    p = O.as_leaves({k: to_t(a, dtype) for k, a in sub(g, "state.").items()})
    x = to_t(g["x"], dtype)
    y = O.model_forward(x, p, c, True, start=start)
    (y * to_t(g["w"], dtype)).sum().backward()
    assert rel_err(y, g["f64.y"]) <= 1e-12
    ours = {k: a.grad for k, a in p.items() if a.requires_grad}
    check_grads(ours, sub(g, "f64.grad."), 1e-9, name)
    with torch.no_grad():
        y_eval = O.model_forward(x, p, c, False, start=start)
    assert rel_err(y_eval, g["f64.y_eval"]) <= 1e-12


@pytest.mark.skipif(not ref_loader.available(), reason="/root/reference not present (GPU box)")
def test_oracle_vs_live_reference_ntu_shape():
    R = ref_loader.load()
    graph = R["Graph"](R["ntu"].skeleton_edges, center_joint=R["ntu"].center_joint)
    adj = R["GraphPartitionStrategy"]().get_adjacency_matrix_array(graph)
    assert np.array_equal(adj, O.partition_adjacency(R["ntu"].skeleton_edges))
    shape = (2, 40, 25, 3)
    state = O.init_state(adj, shape, 60, start=16, seed=5, loud=True)
    ref = R["agcn"].Model(shape, 60, graph, start_feature_size=16)
    ref.load_state_dict(state, strict=True)            # key-for-key compatible
    ref.train()
    x = torch.randn(2, *shape, generator=torch.Generator().manual_seed(9))
    w = torch.randn(2, 60, generator=torch.Generator().manual_seed(10))
    y_ref = ref(x)
    (y_ref * w).sum().backward()
    p = O.as_leaves(state)
    y = O.model_forward(x, p, 3, True, start=16)
    (y * w).sum().backward()
    assert rel_err(y, y_ref) <= 1e-6
    check_grads({k: a.grad for k, a in p.items() if a.requires_grad}, {k: a.grad for k, a in ref.named_parameters()}, 1e-4, "live")
