"""Caller parity: the reference's own multimodal wrappers and model dispatcher, UNMODIFIED, on top
of the B200 unit installed by ``fusion_gcn_b200.dropin``.  Needs the reference checkout (/root/reference, build
container only); on CPU the kernels are replaced by the TEST-ONLY torch stage oracle, so what is checked here is the
drop-in boundary: class rebinding, constructor / forward signatures, state-dict keys, graph objects, input layouts."""
import os
import sys

import numpy as np
import pytest
import torch

from helpers import check_grads, rel_err
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference checkout not present")


@pytest.fixture
def dropin(torch_stage_backend):
    from fusion_gcn_b200 import dropin as D
    D.install_shims(ref_loader.REFERENCE_ROOT)
    yield D
    D.uninstall()


def _reference_then_dropin(D, build, x, w):
    """Builds the model with the reference classes, then again with the drop-in installed and the same weights;
    returns (reference fp64 outputs/grads, drop-in model outputs/grads, models)."""
    torch.manual_seed(0)
    ref = build()
    for name, prm in ref.named_parameters():                     # loud values
        if name.endswith("bn.weight") or name.endswith("down.1.weight"):
            prm.data.uniform_(0.5, 1.5)
        if "adj_b" in name or name.endswith(".PA"):
            prm.data.normal_(0, 0.1)
    state = {k: v.clone() for k, v in ref.state_dict().items()}
    ref64 = build().double()
    ref64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in state.items()})
    ref64.train()
    xd = {k: v.double() for k, v in x.items()} if isinstance(x, dict) else x.double()
    y_ref = ref64(xd)
    (y_ref * w.double()).sum().backward()
    D.install(ref_loader.REFERENCE_ROOT)
    ours = build()
    missing = ours.load_state_dict(state, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ours.train()
    y = ours(x)
    (y * w).sum().backward()
    return ref64, y_ref, ours, y


def _graph(edges, center):
    from util.graph import Graph
    return Graph(edges, center_joint=center)


def test_rebinding_and_restore(dropin):
    import importlib
    ref_m = importlib.import_module("models.mmargcn.agcn")
    original = ref_m.SpatialTemporalConv
    dropin.install(ref_loader.REFERENCE_ROOT)
    from fusion_gcn_b200 import modules, modules_original
    assert ref_m.SpatialTemporalConv is modules.SpatialTemporalConv and ref_m.Model is modules.Model
    ref_o = importlib.import_module("models.agcn.agcn")
    assert ref_o.TCN_GCN_unit is modules_original.TCN_GCN_unit and ref_o.Model is modules_original.Model
    from util.dynamic_import import import_model                  # the reference's own discovery (session.py:50)
    assert import_model("agcn") is modules_original.Model
    dropin.uninstall()
    assert ref_m.SpatialTemporalConv is original
    with pytest.raises(FileNotFoundError):
        dropin.install("/nonexistent")


def test_skeleton_imu_spatial_fusion_wrapper(dropin):
    """config C3: mmargcn.Model(mode='skeleton_imu_spatial_fusion') -> COCO-18 + 4 IMU joints (V = 22), M = 2."""
    import datasets.mmact.constants as mm
    from models.mmargcn import mmargcn
    shape = {"skeleton": (2, 12, 22, 3)}
    build = lambda: mmargcn.Model(shape, 7, _graph(mm.skeleton_edges, mm.center_joint), mode="skeleton_imu_spatial_fusion",   # noqa: E731
                                  num_imu_joints=4, imu_enhanced_mode="append_center", interconnect_imu_joints=True,
                                  center_joint=mm.center_joint, num_layers=3)
    g = torch.Generator().manual_seed(1)
    x, w = torch.randn(3, 2, 12, 22, 3, generator=g), torch.randn(3, 7, generator=g)
    ref64, y_ref, ours, y = _reference_then_dropin(dropin, build, x, w)
    from fusion_gcn_b200 import modules
    assert isinstance(ours._model.agcn, modules.Model)
    assert rel_err(y, y_ref) <= 1e-4
    check_grads({k: p.grad for k, p in ours.named_parameters()}, {k: p.grad for k, p in ref64.named_parameters()}, 1e-4, "imu spatial fusion")


def test_skeleton_imu_channel_fusion_wrapper(dropin):
    """early fusion by channel concatenation: dict input, C = 3 + 6 = 9 (early_fusion_models.py:25-47)."""
    import datasets.utd_mhad.constants as utd
    from models.mmargcn import mmargcn
    shape = {"skeleton": (1, 10, 20, 3), "inertial": (10, 6)}
    build = lambda: mmargcn.Model(shape, 5, _graph(utd.skeleton_edges, utd.center_joint), mode="skeleton_imu_channel_fusion", num_layers=2)   # noqa: E731
    g = torch.Generator().manual_seed(2)
    x = {"skeleton": torch.randn(2, 1, 10, 20, 3, generator=g), "inertial": torch.randn(2, 10, 6, generator=g)}
    w = torch.randn(2, 5, generator=g)
    ref64, y_ref, ours, y = _reference_then_dropin(dropin, build, x, w)
    assert rel_err(y, y_ref) <= 1e-4
    check_grads({k: p.grad for k, p in ours.named_parameters()}, {k: p.grad for k, p in ref64.named_parameters()}, 1e-4, "imu channel fusion")


def test_rgb_patch_features_wrapper(dropin):
    """rgb_patch_features: per-joint 512-d embeddings as channels (rgb_feature_models.py:12-27); reduced to 32-d here."""
    import datasets.utd_mhad.constants as utd
    from models.mmargcn import mmargcn
    shape = {"rgb": (1, 8, 20, 32)}
    build = lambda: mmargcn.Model(shape, 5, _graph(utd.skeleton_edges, utd.center_joint), mode="rgb_patch_features", num_layers=2)   # noqa: E731
    g = torch.Generator().manual_seed(3)
    x, w = torch.randn(2, 1, 8, 20, 32, generator=g), torch.randn(2, 5, generator=g)
    ref64, y_ref, ours, y = _reference_then_dropin(dropin, build, x, w)
    assert rel_err(y, y_ref) <= 1e-4
    check_grads({k: p.grad for k, p in ours.named_parameters()}, {k: p.grad for k, p in ref64.named_parameters()}, 1e-4, "rgb patch features")


def test_unsupported_mode_error_is_the_references(dropin):
    import datasets.utd_mhad.constants as utd
    from models.mmargcn import mmargcn
    dropin.install(ref_loader.REFERENCE_ROOT)
    with pytest.raises(ValueError, match="Unsupported mode"):
        mmargcn.Model({"skeleton": (1, 8, 20, 3)}, 5, _graph(utd.skeleton_edges, utd.center_joint), mode="nope")


def test_imu_gcn_wrapper_runs_on_the_1d_graph_convolutions(dropin, monkeypatch):
    """mmargcn.Model(mode='imu_gcn') -> ImuGCN -> GCN(gc_model='agcn' | 'stgcn') (imu_feature_models.py:63-102, gcn.py:18-83):
    the reference's GCN stack, unmodified, on top of fusion_gcn_b200.graphconv (40-node graph > 32: the large-V path)."""
    import fusion_gcn_b200.modules as MM
    from fusion_gcn_b200 import graphconv
    from models.mmargcn import mmargcn
    monkeypatch.setattr(MM, "_prep", lambda t: t.contiguous())
    for gc_model in ("agcn", "stgcn"):
        shape = {"inertial": (10, 4)}
        build = lambda: mmargcn.Model(shape, 5, None, mode="imu_gcn", gc_model=gc_model, num_layers=3, inner_feature_dim=16)   # noqa: E731
        torch.manual_seed(0)
        ref = build().double()
        state = {k: v.clone() for k, v in ref.state_dict().items()}
        g = torch.Generator().manual_seed(4)
        x, w = torch.randn(3, 10, 4, generator=g, dtype=torch.float64), torch.randn(3, 5, generator=g, dtype=torch.float64)
        y_ref = ref(x)
        (y_ref * w).sum().backward()
        dropin.install(ref_loader.REFERENCE_ROOT)
        ours = build().double()
        ours.load_state_dict(state, strict=True)
        assert isinstance(ours._model.gcn.gc1, (graphconv.AGCNGraphConvolution, graphconv.STGCNGraphConvolution))
        y = ours(x)
        (y * w).sum().backward()
        assert rel_err(y, y_ref) <= 1e-9
        for (k, a), (_, b) in zip(ref.named_parameters(), ours.named_parameters()):
            assert float((a.grad - b.grad).abs().max()) <= 1e-8 * max(1.0, float(a.grad.abs().max())), k
        dropin.uninstall()
