"""Debug aid: run the smoke model with the fused-statistics forward and with the plain one in ONE process and report the
first backward quantities that differ."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fusion_gcn_b200 import graph as G, modules as M, functional as FN, ops as K
from oracle import agcn_oracle as O

fused = FN._conv_bn
def plain(x, w, bias, gamma, beta, buf, training, prec, **kw):
    y = K.conv_fwd(x, w, bias, precision=prec, **kw)
    return y, FN._bn_forward(y, gamma, beta, buf, training)

log = []
orig_t, orig_g = FN.tcn_backward, FN.gcn_backward
def rec_t(d_out, ctx, *a, **k):
    r = orig_t(d_out, ctx, *a, **k)
    log.append(("tcn.d_out_in", d_out.clone())); log.append(("tcn.d_o", None if r[0] is None else r[0].clone()))
    log.append(("tcn.d_xres", None if r[1] is None else r[1].clone()))
    for kk, vv in r[2].items():
        if vv is not None: log.append(("tcn.g." + kk, vv.clone()))
    return r
def rec_g(d_o, ctx, *a, **k):
    r = orig_g(d_o, ctx, *a, **k)
    log.append(("gcn.dx", None if r[0] is None else r[0].clone()))
    for kk, vv in r[1].items():
        if isinstance(vv, list):
            for i, t in enumerate(vv): log.append((f"gcn.g.{kk}.{i}", t.clone()))
        elif vv is not None: log.append(("gcn.g." + kk, vv.clone()))
    return r
FN.tcn_backward, FN.gcn_backward = rec_t, rec_g

start = int(sys.argv[1]) if len(sys.argv) > 1 else 16
shape, n, ncls = (2, 32, 25, 3), 2, 60
graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, seed=1, loud=True)
gen = torch.Generator().manual_seed(2)
x = torch.randn(n, *shape, generator=gen).cuda()
w = torch.randn(n, ncls, generator=gen).cuda()
runs = {}
for name, fn in (("plain", plain), ("fused", fused), ("plain2", plain), ("fused2", fused)):
    FN._conv_bn = fn
    model = M.Model(shape, ncls, graph, start_feature_size=start)
    model.load_state_dict(state, strict=True)
    model.cuda().train()
    log.clear()
    y = model(x)
    (y * w).sum().backward()
    torch.cuda.synchronize()
    runs[name] = list(log)
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
for other in ("plain2", "fused", "fused2"):
    print("==", other, "vs plain")
    shown = 0
    for (k, a), (k2, b) in zip(runs["plain"], runs[other]):
        assert k == k2
        if a is None: continue
        r = rel(b, a)
        if r > 1e-4 and "bias" not in k and not k.endswith((".bd.0", ".bd.1", ".bd.2", ".bt", ".br", "down_b")) and ".ba." not in k:
            print(f"   {k:20s} {tuple(a.shape)} rel diff {r:.3e}")
            shown += 1
            if shown >= 12: break
    if not shown: print("   all backward quantities agree to 1e-4")
