"""Host-side logic on CPU: the autograd composition in fusion_gcn_b200.functional (forward order, every
backward formula, parameter packing), the drop-in modules' state-dict compatibility and layouts.
The kernels are replaced by oracle/stages.py through the TEST-ONLY ``torch_stage_backend`` fixture."""
import numpy as np
import pytest
import torch

from helpers import (MODEL_FIXTURES, RESIDUAL_KINDS, UNIT_FIXTURES, check_grads, eval_mode_gradient_case, load_golden, recompute_case, second_backward_case, rel_err, stat_err, sub,
                     to_t)
from fusion_gcn_b200 import graph as G
from fusion_gcn_b200 import modules as M
from fusion_gcn_b200 import modules_original as MO


def test_graph_builder_matches_reference_golden():
    g = load_golden("adjacency")
    for name, edges in (("ntu", G.NTU_EDGES), ("utd", G.UTD_EDGES), ("mmact", G.MMACT_EDGES)):
        assert np.array_equal(np.unique(np.asarray(edges), axis=0), np.unique(g[name + "_edges"], axis=0))
        assert np.array_equal(G.partition_adjacency(edges), g[name + "_adj"]), name
    base = G.SkeletonGraph(G.MMACT_EDGES, center_joint=G.MMACT_CENTER)
    for inter in (False, True):
        fused = G.imu_fusion_graph(base, 4, "append_center", interconnect=inter)
        assert fused.num_vertices == 22
        assert np.array_equal(G.adjacency_from_graph(fused), g["mmact_imu4_" + ("inter" if inter else "plain") + "_adj"])
    with pytest.raises(NotImplementedError):
        G.partition_adjacency(G.NTU_EDGES, strategy="distance")
    with pytest.raises(ValueError):
        G.imu_fusion_graph(base, 2, "append_left")


def test_cuda_path_has_no_cpu_fallback():
    """Without the test backend the modules must refuse CPU tensors loudly."""
    unit = M.SpatialTemporalConv(4, 8, G.partition_adjacency(G.UTD_EDGES))
    with pytest.raises(RuntimeError, match="CUDA|libagcn_b200"):
        unit(torch.randn(1, 4, 6, 20))


@pytest.mark.parametrize("name", UNIT_FIXTURES)
def test_unit_module_vs_golden(name, torch_stage_backend):
    g = load_golden("unit_" + name)
    cin, cout, stride, res = [int(v) for v in g["meta"]]
    state = {k: to_t(v) for k, v in sub(g, "state.").items()}
    adj = state["gcn1.adj_a"].numpy().astype(np.float64)
    unit = M.SpatialTemporalConv(cin, cout, adj, stride=stride, residual=(res != 0))
    assert sorted(unit.state_dict().keys()) == sorted(state.keys())
    unit.load_state_dict(state, strict=True)
    unit.train()
    x = to_t(g["x"]).requires_grad_(True)            # reference layout (N', C, T, V)
    y = unit(x)
    assert y.shape == g["f64.y"].shape and y.is_contiguous()
    (y * to_t(g["w"])).sum().backward()
    assert rel_err(y, g["f64.y"]) <= 2e-5
    assert rel_err(x.grad, g["f64.dx"]) <= 1e-4
    for k in range(3):
        assert rel_err(unit.gcn1.adj_c[k], g[f"f64.adj_c.{k}"]) <= 1e-5
        assert not unit.gcn1.adj_c[k].requires_grad
    check_grads({k: p.grad for k, p in unit.named_parameters()}, sub(g, "f64.grad."), 1e-4, name)
    for k, v in sub(g, "f64.after.").items():
        assert stat_err(unit.state_dict()[k], v) <= 1e-5, k


@pytest.mark.parametrize("name", MODEL_FIXTURES)
def test_model_module_vs_golden(name, torch_stage_backend):
    g = load_golden("model_" + name)
    m, t, v, c, ncls, start = [int(a) for a in g["meta"]]
    state = {k: to_t(a) for k, a in sub(g, "state.").items()}
    graph = G.SkeletonGraph(G.UTD_EDGES if v == 20 else G.NTU_EDGES)
    model = M.Model((m, t, v, c), ncls, graph, start_feature_size=start)
    assert sorted(model.state_dict().keys()) == sorted(state.keys())
    model.load_state_dict(state, strict=True)
    model.train()
    y = model(to_t(g["x"]))
    (y * to_t(g["w"])).sum().backward()
    assert rel_err(y, g["f64.y"]) <= 1e-4
    check_grads({k: p.grad for k, p in model.named_parameters()}, sub(g, "f64.grad."), 1e-4 if "default" not in name else 2e-4, name)
    for k, v_ in sub(g, "f64.after.").items():
        assert stat_err(model.state_dict()[k], v_) <= 1e-4, k
    assert int(model.data_bn.num_batches_tracked) == 1 and int(model.l3.tcn1.bn.num_batches_tracked) == 1
    model.eval()
    with torch.no_grad():
        assert rel_err(model(to_t(g["x"])), g["f64.y_eval"]) <= 1e-4


def test_standalone_halves_and_kwargs(torch_stage_backend):
    torch.manual_seed(0)
    adj = G.partition_adjacency(G.UTD_EDGES)
    tcn = M.TemporalConv(8, 12, kernel_size=9, stride=2)
    ref_tcn = torch.nn.Sequential(torch.nn.Conv2d(8, 12, (9, 1), padding=(4, 0), stride=(2, 1)), torch.nn.BatchNorm2d(12))
    ref_tcn[0].load_state_dict(tcn.conv.state_dict())
    x = torch.randn(2, 8, 11, 20)
    assert rel_err(tcn(x), ref_tcn(x)) <= 1e-5                     # no activation inside TemporalConv
    gcn = M.SpatialGraphConv(8, 8, adj)
    y = gcn(x)
    assert y.shape == x.shape and float(y.min()) >= 0.0
    assert all(a.shape == (2, 20, 20) for a in gcn.adj_c)
    assert torch.allclose(gcn.adj_c[0].sum(dim=-2), torch.ones(2, 20), atol=1e-5)      # columns sum to one
    model = M.Model((1, 12, 20, 3), 5, G.SkeletonGraph(G.UTD_EDGES), num_layers=3, start_feature_size=8, without_fc=True, dropout=0.25)
    assert model.fc is None and model.out_channels == 8
    assert isinstance(model.l1, torch.nn.Dropout) and isinstance(model.l2, M.SpatialTemporalConv)      # agcn.py:166-172 naming
    assert model(torch.randn(2, 1, 12, 20, 3)).shape == (2, 8)
    with pytest.raises(ValueError):
        M.SpatialGraphConv(8, 8, adj[:1], num_subsets=1)


def test_original_variant_state_dict_and_parity(torch_stage_backend):
    """models/agcn/agcn.py: layers l1..l10, parameter PA, adjacency not in the state dict, dict data_shape."""
    g = load_golden("model_ntu_s8_m2")
    m, t, v, c, ncls, start = [int(a) for a in g["meta"]]
    graph = G.SkeletonGraph(G.NTU_EDGES)
    model = MO.Model({"skeleton": (m, t, v, c)}, ncls, graph, start_feature_size=start)
    keys = list(model.state_dict().keys())
    assert "l1.gcn1.PA" in keys and "l10.tcn1.conv.weight" in keys and not any("adj_a" in k or k.startswith("l0.") for k in keys)
    state = {}
    for k, a in sub(g, "state.").items():
        if k.endswith("adj_a"):
            continue
        if k[0] == "l" and k[1].isdigit():
            idx, rest = k[1:].split(".", 1)
            k = f"l{int(idx) + 1}.{rest}"
        state[k.replace("adj_b", "PA")] = to_t(a)
    model.load_state_dict(state, strict=True)
    model.train()
    y = model(to_t(g["x"]))
    assert rel_err(y, g["f64.y"]) <= 1e-4
    assert isinstance(model.l1, MO.TCN_GCN_unit) and isinstance(model.l1.gcn1, MO.unit_gcn) and isinstance(model.l1.tcn1, MO.unit_tcn)


def test_state_dict_order_matches_live_reference():
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("/root/reference not present (GPU box)")
    R = ref_loader.load()
    rg = R["Graph"](R["ntu"].skeleton_edges, center_joint=R["ntu"].center_joint)
    ref = R["agcn"].Model((2, 16, 25, 3), 60, rg, start_feature_size=8)
    ours = M.Model((2, 16, 25, 3), 60, rg, start_feature_size=8)          # accepts the reference's own Graph object
    assert list(ours.state_dict().keys()) == list(ref.state_dict().keys())
    assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    assert len(ours.state_dict()) == 362 and sum(p.numel() for p in ours.parameters()) == sum(p.numel() for p in ref.parameters())
    # same initial distributions: identical tensors under the same seed
    torch.manual_seed(4)
    a = R["agcn"].SpatialTemporalConv(8, 16, np.asarray(ours.l0.gcn1.adj_a.numpy(), dtype=np.float64), stride=2)
    torch.manual_seed(4)
    b = M.SpatialTemporalConv(8, 16, np.asarray(ours.l0.gcn1.adj_a.numpy(), dtype=np.float64), stride=2)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb), ka


def test_bench_reference_arm_line_schema():
    """`bench.py --impl reference` (the tier's CPU arm: the unmodified reference from /root/reference or its byte copy under
    baseline/_ref -- the oracle port only when neither exists -- on the host cores, bounded sample) prints one JSON line with
    the contract's keys; runs without a GPU."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--workload", "utd"],
                         capture_output=True, text=True, timeout=600, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "sequences/s" and line["value"] > 0 and line["higher_is_better"] is True
    from oracle import ref_loader
    assert line["cpu_baseline"]["kind"] == ("reference" if ref_loader.available() else "port") and line["cpu_baseline"]["cores"] >= 1 and "sample" in line["cpu_baseline"]
    assert line["e2e"] == {"value": line["value"], "unit": "sequences/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


@pytest.mark.parametrize("cin,cout,stride,residual,v", [(8, 8, 1, True, 25), (8, 16, 2, True, 22), (3, 8, 1, False, 20)])
def test_unit_parity_harness_on_cpu(torch_stage_backend, cin, cout, stride, residual, v):
    """The BASELINE-shape parity harness (tests/unit_parity.py: plain fp64 forward, ReLU brackets equal except at ties,
    gradients on the same linear piece at 1e-4) run on CPU over the torch stage backend, so the harness itself is checked
    without a GPU."""
    import unit_parity as UP
    from fusion_gcn_b200 import graph as G, modules as M
    unit = UP.baseline_unit(M, G, cin, cout, stride, residual, v, seed=3)
    x, w = UP.unit_inputs(3, cin, cout, 21, v, stride, seed=4)
    err = UP.run_unit_parity(unit, x, w, "cpu")
    assert err["y"] <= 1e-5 and err["dx"] <= 1e-5


def test_model_level_cases_of_the_gpu_suite_on_cpu(torch_stage_backend):
    """The original-variant and dropout model cases of tests/test_gpu_baseline_shapes.py (strict load of reference-keyed
    states, mask capture and replay in the oracle) over the torch stage backend."""
    import test_gpu_baseline_shapes as T
    T.run_original_variant("cpu")
    T.run_dropout_model("cpu")


def test_model_parity_harness_on_cpu(torch_stage_backend):
    """tests/unit_parity.py::run_model_parity (the model-level 1e-4 contract of the GPU suite) over the torch stage backend."""
    import test_gpu_unit as T
    from fusion_gcn_b200 import graph as G, modules as M
    err = T.seeded_model_case(M, G, (1, 20, 20, 9), "utd", 16, 3, "fp32", "cpu")
    assert err["y"] <= 1e-5


def test_fused_head_and_pooled_tail_match_the_separate_ops(torch_stage_backend):
    """Model.loss (fc + cross-entropy as one node) and the pooled tail of the last unit against nn.CrossEntropyLoss()(model(x), y)
    built from the separate pool / linear ops: same loss, logits and gradients (host composition, torch stage backend)."""
    import copy
    from fusion_gcn_b200 import functional as FN, graph as G, modules as M
    from fusion_gcn_b200.graphed import loss_and_logits
    torch.manual_seed(3)
    model = M.Model((2, 12, 25, 3), 11, G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER), start_feature_size=8).double().train()
    ref = copy.deepcopy(model)
    x = torch.randn(3, 2, 12, 25, 3, dtype=torch.float64)
    y = torch.tensor([1, 7, 10])
    import fusion_gcn_b200.modules as MM
    orig = MM._prep
    MM._prep = lambda t: t.contiguous()                  # keep fp64 through the stage backend
    try:
        loss, logits = loss_and_logits(model, torch.nn.CrossEntropyLoss(), x, y)
        loss.backward()
        h = ref.features_cl(x)                                                    # unfused: feature map -> PoolFn -> LinearFn -> torch CE
        feat = FN.PoolFn.apply(h, 3)
        logits_ref = FN.LinearFn.apply(feat, ref.fc.weight, ref.fc.bias, 0)
        loss_ref = torch.nn.functional.cross_entropy(logits_ref, y)
        loss_ref.backward()
    finally:
        MM._prep = orig
    assert abs(float(loss) - float(loss_ref)) <= 1e-12 and torch.allclose(logits, logits_ref, atol=1e-12)
    for (k, a), (_, b) in zip(model.named_parameters(), ref.named_parameters()):
        assert torch.allclose(a.grad, b.grad, rtol=1e-9, atol=1e-12), k


def test_wide_odd_channel_counts_are_zero_padded_for_the_first_unit(torch_stage_backend):
    """C = 70 stands in for the 515-channel skeleton + RGB-patch fusion (early_fusion_models.py:53-60): Model pads the first
    unit's input (and, through F.pad, its weights) to a multiple of 32; logits and every gradient (in the parameters' own shapes)
    still match the oracle."""
    import test_gpu_unit as T
    from fusion_gcn_b200 import graph as G, modules as M
    assert M._padded_channels(515) == 544 and M._padded_channels(512) == 512 and M._padded_channels(9) == 9
    err = T.seeded_model_case(M, G, (1, 12, 20, 70), "utd", 8, 3, "fp32", "cpu")
    assert err["y"] <= 1e-5


def test_eval_mode_gradients_match_the_oracle(torch_stage_backend):
    eval_mode_gradient_case(M, G, "cpu", 1e-4)



def test_recompute_policy_gives_identical_gradients(torch_stage_backend):
    recompute_case(M, G, "cpu")


def test_eval_no_grad_takes_the_fused_nothing_saved_path(torch_stage_backend, monkeypatch):
    """model.eval() under torch.no_grad() (session.py:188-194) must run the fused tails (conv_fwd_post), not the unfused pipeline that a
    backward could follow: ctx.needs_input_grad ignores the grad mode, so the caller's grad mode is recorded in the UnitSpec."""
    import collections
    import fusion_gcn_b200.functional as FN
    calls = collections.Counter()

    class Counting:
        def __getattr__(self, name):
            attr = getattr(torch_stage_backend, name)
            if not callable(attr):
                return attr

            def wrapped(*a, **kw):
                calls[name] += 1
                return attr(*a, **kw)
            return wrapped

    monkeypatch.setattr(FN, "K", Counting())
    model = M.Model((2, 16, 25, 3), 60, G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER), start_feature_size=16).eval()
    x = torch.randn(2, 2, 16, 25, 3)
    with torch.no_grad():
        y0 = model(x)
    assert calls["conv_fwd_post"] == 25 and calls["bn_apply"] == 2, calls          # 2 = data_bn of the two bodies
    calls.clear()
    y1 = model(x)                                   # grad mode on, parameters require grad: a backward may follow
    assert calls["conv_fwd_post"] == 0 and calls["bn_apply"] == 22, calls
    assert rel_err(y0, y1.detach()) <= 1e-6


@pytest.mark.parametrize("training", [True, False])
def test_second_backward_through_the_same_forward(torch_stage_backend, training):
    unit = M.SpatialTemporalConv(8, 16, G.partition_adjacency(G.UTD_EDGES), stride=2)
    unit.train(training)
    second_backward_case(unit, torch.randn(2, 8, 12, 20))
