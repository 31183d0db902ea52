"""Debug aid: for every Conv->BN pair of the smoke model, compare the fused-epilogue statistics with bn_stats on the same y."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fusion_gcn_b200 import graph as G, modules as M, functional as FN, ops as K
from oracle import agcn_oracle as O

orig = FN._conv_bn
def probe(x, w, bias, gamma, beta, buf, training, prec, **kw):
    rm, rv = buf.running_mean.clone(), buf.running_var.clone()
    y, st = orig(x, w, bias, gamma, beta, buf, training, prec, **kw)
    ref = K.bn_stats(y, gamma, beta, rm, rv, None, 0.1, 1e-5, True)
    d = [float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(st, ref)]
    print(f"x{tuple(x.shape)} w{tuple(w.shape)} kw={kw} rel diffs scale/shift/mean/invstd: " + " ".join(f"{v:.2e}" for v in d), flush=True)
    return y, st
FN._conv_bn = probe
start = 16
shape, n, ncls = (2, 32, 25, 3), 2, 60
graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, seed=1, loud=True)
gen = torch.Generator().manual_seed(2)
x = torch.randn(n, *shape, generator=gen)
model = M.Model(shape, ncls, graph, start_feature_size=start)
model.load_state_dict(state, strict=True)
model.cuda().train()
y = model(x.cuda())
torch.cuda.synchronize()
