"""-m gpu: parity at the shapes that matter (VERDICT r1 "do this" item 1).

(a) ONE SpatialTemporalConv at each BASELINE layer shape (3->64, 64->64, 64->128 s2, 128->128, 128->256 s2, 256->256;
    T = 300 / 150 / 75; V = 25 and V = 22; N' = 16) in every parity mode, against the fp64 oracle at 1e-4 on y, dx and
    EVERY gradient with loud initialisation and no noise-scaled slack (tests/unit_parity.py explains how ReLU ties are
    handled: brackets must agree except within 1e-5 of zero, gradients are compared on the same linear piece);
(b) the original-variant Model (models/agcn/agcn.py) on the GPU with a strict load of a reference-keyed state dict;
(c) one optimisation step under torch.autocast + GradScaler (procedures/step.py:55-78) and one with dropout = 0.5;
(d) the real sequence lengths T = 128 (UTD-MHAD) and T = 515 (MMAct, odd through a stride-2 layer).
Reference: torch_src/models/mmargcn/agcn.py:118-200, torch_src/models/agcn/agcn.py:116-191."""
import numpy as np
import pytest
import torch

import unit_parity as UP
from helpers import ZERO_GRAD, check_grads, rel_err
from oracle import agcn_oracle as O

pytestmark = pytest.mark.gpu
from fusion_gcn_b200.modules import _PRECISIONS
PARITY_MODES = [m for m in ("fp32", "bf16x3") if m in _PRECISIONS]      # every mode that claims the 1e-4 contract


@pytest.fixture(scope="module")
def pkg():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    import fusion_gcn_b200
    from fusion_gcn_b200 import capi
    capi.lib()
    return fusion_gcn_b200


BASELINE_LAYERS = [        # cin, cout, stride, residual, T   (mmargcn/agcn.py:152-163 at NTU T = 300)
    (3, 64, 1, False, 300), (64, 64, 1, True, 300), (64, 128, 2, True, 300), (128, 128, 1, True, 150),
    (128, 256, 2, True, 150), (256, 256, 1, True, 75)]


@pytest.mark.parametrize("precision", PARITY_MODES)
@pytest.mark.parametrize("v", [25, 22])
@pytest.mark.parametrize("cin,cout,stride,residual,t", BASELINE_LAYERS)
def test_unit_at_baseline_layer_shape(pkg, cin, cout, stride, residual, t, v, precision):
    from fusion_gcn_b200 import graph as G, modules as M
    unit = UP.baseline_unit(M, G, cin, cout, stride, residual, v, seed=11)
    M.set_precision(unit, precision)
    x, w = UP.unit_inputs(16, cin, cout, t, v, stride, seed=12)
    err = UP.run_unit_parity(unit, x, w, "cuda")
    print(f"[{precision}] {cin}->{cout} s{stride} T={t} V={v}: y {err['y']:.2e} dx {err['dx']:.2e} worst grad {err['worst_grad'][1]:.2e} "
          f"({err['worst_grad'][0]}), ReLU ties {err['relu_ties']}")


@pytest.mark.parametrize("precision", PARITY_MODES)
@pytest.mark.parametrize("cin,cout,stride,t,v,nb", [(64, 64, 1, 128, 20, 4), (64, 128, 2, 515, 22, 4), (128, 128, 1, 258, 22, 4),
                                                   (128, 256, 2, 129, 22, 4)])
def test_unit_at_real_sequence_lengths(pkg, cin, cout, stride, t, v, nb, precision):
    """UTD-MHAD skeletons have T = 128, MMAct T = 515 -> 258 -> 129 through the two stride-2 layers."""
    from fusion_gcn_b200 import graph as G, modules as M
    unit = UP.baseline_unit(M, G, cin, cout, stride, True, v, seed=21)
    M.set_precision(unit, precision)
    x, w = UP.unit_inputs(nb, cin, cout, t, v, stride, seed=22)
    UP.run_unit_parity(unit, x, w, "cuda")


def test_original_variant_model_on_gpu(pkg):
    """models/agcn/agcn.py drop-in: reference-keyed state (PA, l1..l10, no adjacency entry) loads strictly, logits and
    gradients match the oracle run of the same variant."""
    run_original_variant("cuda")


def run_original_variant(device):
    from fusion_gcn_b200 import graph as G, modules_original as MO
    graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
    adj = G.adjacency_from_graph(graph)
    shape, ncls, start, n = (2, 32, 25, 3), 60, 16, 2
    state = O.init_state(adj, shape, ncls, start=start, seed=5, loud=True, variant="original")
    model = MO.Model({"skeleton": shape}, ncls, graph, start_feature_size=start, mode="ignored_like_the_reference")
    model.load_state_dict(state, strict=True)
    names = [k for k, _ in model.named_parameters()]
    assert names.index("l1.gcn1.PA") < names.index("l1.gcn1.conv_a.0.weight")      # PA first inside the unit, as in the reference
    gen = torch.Generator().manual_seed(6)
    x = torch.randn(n, *shape, generator=gen)
    w = torch.randn(n, ncls, generator=gen)
    # logits + every gradient at 1e-4 against the oracle run of the same variant (ReLU brackets pinned as in tests/unit_parity.py)
    err = UP.run_model_parity(model, state, x, w, device, 3, start, variant="original", adj_a=adj)
    print(f"original variant: logits {err['y']:.2e}, worst grad {err['worst_grad']}, ReLU ties {err['relu_ties']}")


def _small_model(M, G, dropout=0.0, seed=8, device="cuda"):
    graph = G.SkeletonGraph(G.UTD_EDGES, center_joint=G.UTD_CENTER)
    shape, ncls, start = (1, 24, 20, 3), 27, 16
    state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, seed=seed, loud=True)
    model = M.Model(shape, ncls, graph, start_feature_size=start, dropout=dropout)
    if dropout > 0:      # the reference renumbers: units take the even names, nn.Dropout the odd ones (agcn.py:166-172)
        state = {(f"l{2 * int(k[1:k.index('.')])}{k[k.index('.'):]}" if k[0] == "l" and k[1].isdigit() else k): v for k, v in state.items()}
    model.load_state_dict(state, strict=True)
    return model.to(device).train(), state, shape, ncls, start


def test_autocast_and_gradscaler_step(pkg):
    """MixedPrecisionStep (procedures/step.py:55-78): forward + loss under torch.autocast, GradScaler-scaled backward,
    unscale, optimizer step.  The kernels compute in fp32 whatever the autocast state is, so the unscaled gradients must
    equal the plain step's and the scaler must see finite values."""
    from fusion_gcn_b200 import graph as G, modules as M
    model, state, shape, ncls, start = _small_model(M, G)
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(4, *shape, generator=gen).cuda()
    lab = torch.randint(0, ncls, (4,), generator=gen).cuda()
    loss_fn = torch.nn.CrossEntropyLoss()
    loss_fn(model(x), lab).backward()
    plain = {k: p.grad.clone() for k, p in model.named_parameters()}
    model.load_state_dict(state, strict=True)          # running statistics back to the start
    model.zero_grad(set_to_none=True)
    opt = torch.optim.SGD(model.parameters(), lr=0.1, momentum=0.9, nesterov=True, weight_decay=1e-4)
    scaler = torch.amp.GradScaler("cuda", init_scale=1024.0)
    with torch.autocast("cuda"):
        y = model(x)
        loss = loss_fn(y, lab)
    assert y.dtype == torch.float32
    scaler.scale(loss).backward()
    scaler.unscale_(opt)
    for k, p in model.named_parameters():
        assert torch.isfinite(p.grad).all(), k
        if not ZERO_GRAD.search(k):
            assert rel_err(p.grad, plain[k]) <= 1e-5, k
    before = model.fc.weight.detach().clone()
    scaler.step(opt)
    scaler.update()
    assert scaler.get_scale() == 1024.0 and not torch.equal(before, model.fc.weight)


def test_dropout_model_against_oracle_with_the_same_masks(pkg):
    """Model(dropout=0.5): nn.Dropout(inplace=True) runs between our units on the channels-last tensors.  The masks torch
    draws are captured with hooks and replayed in the fp64 oracle (reference layout), so logits and gradients are compared
    on identical masks."""
    run_dropout_model("cuda")


def run_dropout_model(device):
    from fusion_gcn_b200 import graph as G, modules as M
    model, state, shape, ncls, start = _small_model(M, G, dropout=0.5, device=device)
    drops = [m for m in model.layers if isinstance(m, torch.nn.Dropout)]
    assert len(drops) == 9 and all(d.inplace for d in drops)
    gen = torch.Generator().manual_seed(10)
    x = torch.randn(3, *shape, generator=gen)
    w = torch.randn(3, ncls, generator=gen)
    torch.manual_seed(1)
    # logits against the plain fp64 oracle on the same masks; ReLU brackets equal except at ties; gradients on our linear piece, 1e-4
    err = UP.run_model_parity(model, state, x, w, device, 3, start)
    print(f"dropout model: y {err['y']:.2e} worst grad {err['worst_grad'][1]:.2e} ({err['worst_grad'][0]}), ReLU ties {err['relu_ties']}")


def test_second_device_and_foreign_current_device(pkg):
    """the shared-memory opt-in is per device and the launch must follow the tensors' device, not the current one."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from fusion_gcn_b200 import graph as G, modules as M
    unit = UP.baseline_unit(M, G, 64, 64, 1, True, 25, seed=3)
    x, w = UP.unit_inputs(2, 64, 64, 20, 25, 1, seed=4)
    y0 = unit.to("cuda:0")(x.to("cuda:0"))
    torch.cuda.set_device(0)
    y1 = unit.to("cuda:1")(x.to("cuda:1"))            # current device stays cuda:0
    assert torch.equal(y0.cpu(), y1.cpu())


@pytest.mark.parametrize("precision", PARITY_MODES)
def test_rgb_patch_early_fusion_channel_count_on_the_tensor_cores(pkg, precision):
    """BASELINE configs[3]: skeleton + per-joint 512-d RGB patch embeddings as extra channels, C = 3 + 512 = 515
    (early_fusion_models.py:53-60), V = 25.  515 is not a multiple of 4, so Model zero-pads the first unit's input to 544
    channels and the unit runs on the TMA / tcgen05 kernels; logits and all gradients against the fp64 oracle."""
    import test_gpu_unit as T
    from fusion_gcn_b200 import capi, graph as G, modules as M, ops
    ops.start_timing(("*",))
    err = T.seeded_model_case(M, G, (1, 24, 25, 515), "ntu", 64, 2, precision, "cuda", tol=3e-4 if precision == "bf16x3" else 1e-4)
    sigs = ops.stop_timing()
    first = [sig for (name, sig) in sigs if name == "agcn_conv_fwd" and sig[4] in (544, 3 * 544)]
    assert first, "the first unit's convolutions must see the padded 544-channel input"
    print(f"[{precision}] C=515: logits {err['y']:.2e}, worst grad {err['worst_grad']}")
