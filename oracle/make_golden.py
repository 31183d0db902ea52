"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz from the UNMODIFIED
reference (needs /root/reference; run in the build container):

    python oracle/make_golden.py

Every fixture holds the seeded input, the full state dict (reference key names),
and what the reference modules produced for it in fp32 and in fp64 (ground
truth): outputs, the attention matrices adj_c, every parameter gradient for the
upstream gradient stored alongside, and the BN running statistics after the
step.  "Loud" initialisation is used so the adaptive branch matters.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import agcn_oracle as O          # noqa: E402
from oracle import ref_loader                # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _np(t):
    return t.detach().cpu().numpy()


def adjacency_fixtures(R):
    out = {}
    S = R["GraphPartitionStrategy"]()
    for name in ("ntu", "utd", "mmact"):
        c = R[name]
        g = R["Graph"](c.skeleton_edges, center_joint=c.center_joint)
        out[name + "_edges"] = np.asarray(c.skeleton_edges)
        out[name + "_center"] = np.asarray(c.center_joint)
        out[name + "_adj"] = S.get_adjacency_matrix_array(g)
    c = R["mmact"]
    g = R["Graph"](c.skeleton_edges, center_joint=c.center_joint)
    for inter in (False, True):
        g2 = R["fusion"].get_skeleton_imu_fusion_graph(g, "append_center", 4, interconnect_imu_joints=inter)
        out["mmact_imu4_" + ("inter" if inter else "plain") + "_adj"] = S.get_adjacency_matrix_array(g2)
        out["mmact_imu4_" + ("inter" if inter else "plain") + "_edges"] = np.asarray(g2.edges)
    for k, v in out.items():
        assert np.isfinite(v).all(), k
    np.savez_compressed(os.path.join(OUT, "adjacency.npz"), **out)
    return out


def run_module(mod, x, w, dtype):
    mod = mod.to(dtype)
    mod.train()
    x = x.to(dtype).clone().requires_grad_(True)
    y = mod(x)
    (y * w.to(dtype)).sum().backward()
    res = {"y": _np(y), "dx": _np(x.grad)}
    for k, p in mod.named_parameters():
        res["grad." + k] = _np(p.grad)
    for k, b in mod.state_dict().items():
        if "running_" in k:
            res["after." + k] = _np(b)
    return res


def unit_fixture(R, name, adj, cin, cout, stride, residual, n, t, seed):
    ref = R["agcn"]
    v = adj.shape[-1]
    g = torch.Generator().manual_seed(seed)
    state = _unit_state(adj, cin, cout, residual, seed)
    x = torch.randn(n, cin, t, v, generator=g)
    t_out = (t - 1) // stride + 1
    w = torch.randn(n, cout, t_out, v, generator=g)
    save = {"x": _np(x), "w": _np(w), "meta": np.asarray([cin, cout, stride, {"none": 0, "identity": 1, "conv": 2}[residual]])}
    for k, tns in state.items():
        save["state." + k] = _np(tns)
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        mod = ref.SpatialTemporalConv(cin, cout, adj, stride=stride, residual=(residual != "none"))
        mod.load_state_dict(state, strict=True)
        r = run_module(mod, x, w, dtype)
        for k, val in r.items():
            save[tag + "." + k] = val
        for i, pk in enumerate(mod.gcn1.adj_c):
            save[f"{tag}.adj_c.{i}"] = _np(pk)
    np.savez_compressed(os.path.join(OUT, f"unit_{name}.npz"), **save)


def _unit_state(adj, cin, cout, residual, seed):
    """Loud unit state with reference key names (no layer prefix)."""
    import math
    g = torch.Generator().manual_seed(1000 + seed)
    a32 = torch.from_numpy(adj.astype(np.float32))
    p = {}

    def bn(prefix, ch):
        p[prefix + ".weight"] = torch.rand(ch, generator=g) + 0.5
        p[prefix + ".bias"] = torch.rand(ch, generator=g) * 0.4 - 0.2
        p[prefix + ".running_mean"] = torch.zeros(ch)
        p[prefix + ".running_var"] = torch.ones(ch)
        p[prefix + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    def conv(prefix, co, ci_, k=1, std=None):
        std = math.sqrt(2.0 / (co * k)) if std is None else std
        p[prefix + ".weight"] = torch.randn(co, ci_, k, 1, generator=g) * std
        p[prefix + ".bias"] = torch.randn(co, generator=g) * 0.05

    p["gcn1.adj_b"] = torch.randn(a32.shape, generator=g) * 0.1
    p["gcn1.adj_a"] = a32.clone()
    for k in range(3):
        conv(f"gcn1.conv_a.{k}", cout // 4, cin)
        conv(f"gcn1.conv_b.{k}", cout // 4, cin)
        conv(f"gcn1.conv_d.{k}", cout, cin, std=math.sqrt(2.0 / (cout * cin * 3)) * 3)
    if cin != cout:
        conv("gcn1.down.0", cout, cin)
        bn("gcn1.down.1", cout)
    bn("gcn1.bn", cout)
    conv("tcn1.conv", cout, cout, 9)
    bn("tcn1.bn", cout)
    if residual == "conv":
        conv("residual.conv", cout, cin, 1)
        bn("residual.bn", cout)
    return p


def model_fixture(R, name, edges, center, data_shape, num_classes, n, start, seed, loud=True):
    ref = R["agcn"]
    g = R["Graph"](edges, center_joint=center)
    adj = R["GraphPartitionStrategy"]().get_adjacency_matrix_array(g)
    m_, t, v, c = data_shape
    state = O.init_state(adj, data_shape, num_classes, start=start, seed=seed, loud=loud)
    gen = torch.Generator().manual_seed(seed + 77)
    x = torch.randn(n, m_, t, v, c, generator=gen)
    w = torch.randn(n, num_classes, generator=gen)
    save = {"x": _np(x), "w": _np(w), "adj": adj,
            "meta": np.asarray([m_, t, v, c, num_classes, start])}
    for k, tns in state.items():
        save["state." + k] = _np(tns)
    for tag, dtype in (("f32", torch.float32), ("f64", torch.float64)):
        mod = ref.Model(data_shape, num_classes, g, start_feature_size=start)
        mod.load_state_dict(state, strict=True)
        r = run_module(mod, x, w, dtype)
        r.pop("dx")
        for k, val in r.items():
            save[tag + "." + k] = val
        # eval-mode logits with the updated running statistics
        mod.eval()
        with torch.no_grad():
            save[tag + ".y_eval"] = _np(mod(x.to(dtype)))
    np.savez_compressed(os.path.join(OUT, f"model_{name}.npz"), **save)


def main():
    os.makedirs(OUT, exist_ok=True)
    R = ref_loader.load()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    adj = adjacency_fixtures(R)
    # the five structural variants of SpatialTemporalConv, scaled down
    unit_fixture(R, "first_c3_16", adj["ntu_adj"], 3, 16, 1, "none", n=3, t=12, seed=1)
    unit_fixture(R, "same_16_16", adj["ntu_adj"], 16, 16, 1, "identity", n=2, t=12, seed=2)
    unit_fixture(R, "down_16_32_s2", adj["utd_adj"], 16, 32, 2, "conv", n=2, t=13, seed=3)      # odd T
    unit_fixture(R, "same_32_32_v22", adj["mmact_imu4_inter_adj"], 32, 32, 1, "identity", n=2, t=10, seed=4)
    unit_fixture(R, "wide_c9_16", adj["utd_adj"], 9, 16, 1, "conv", n=2, t=9, seed=5)              # Cin != Cout, stride 1
    model_fixture(R, "utd_s8", R["utd"].skeleton_edges, R["utd"].center_joint, (1, 24, 20, 3), 27, n=3, start=8, seed=11)
    model_fixture(R, "ntu_s8_m2", R["ntu"].skeleton_edges, R["ntu"].center_joint, (2, 20, 25, 3), 60, n=2, start=8, seed=12)
    model_fixture(R, "utd_s8_default_init", R["utd"].skeleton_edges, R["utd"].center_joint, (1, 16, 20, 3), 27, n=2,
                  start=8, seed=13, loud=False)
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
