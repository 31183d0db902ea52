"""Debug aid: log every stage call (name, shapes, outputs) of a fwd+bwd with the fused-statistics forward and with the plain
one, in ONE process, and print the first calls whose outputs differ."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fusion_gcn_b200 import graph as G, modules as M, functional as FN, ops
from oracle import agcn_oracle as O

class Rec:
    def __init__(self): self.log = []
    def __getattr__(self, name):
        fn = getattr(ops, name)
        if not callable(fn): return fn
        def wrap(*a, **k):
            r = fn(*a, **k)
            outs = r if isinstance(r, (tuple, list)) else (r,)
            ins = [tuple(t.shape) for t in a if torch.is_tensor(t)]
            self.log.append((name, ins, {kk: vv for kk, vv in k.items() if not torch.is_tensor(vv)}, [o.clone() if torch.is_tensor(o) else None for o in outs], [t.clone() for t in a if torch.is_tensor(t)] if name == 'bn_bwd' else None))
            return r
        return wrap
rec = Rec()
FN.K = rec
fused = FN._conv_bn
def plain(x, w, bias, gamma, beta, buf, training, prec, **kw):
    y = rec.conv_fwd(x, w, bias, precision=prec, **kw)
    return y, FN._bn_forward(y, gamma, beta, buf, training)

start = 16
shape, n, ncls = (2, 32, 25, 3), 2, 60
graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, seed=1, loud=True)
gen = torch.Generator().manual_seed(2)
x = torch.randn(n, *shape, generator=gen).cuda()
w = torch.randn(n, ncls, generator=gen).cuda()
runs = {}
fwd = {}
for name, fn in (("plain", plain), ("fused", fused)):
    FN._conv_bn = fn
    rec.log.clear()
    model = M.Model(shape, ncls, graph, start_feature_size=start)
    model.load_state_dict(state, strict=True)
    model.cuda().train()
    y = model(x)
    torch.cuda.synchronize()
    fwd[name] = list(rec.log)
    rec.log.clear()                       # backward only
    (y * w).sum().backward()
    torch.cuda.synchronize()
    runs[name] = list(rec.log)
def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
shown = 0
print(len(runs["plain"]), len(runs["fused"]))
for i, ((n1, i1, k1, o1, a1), (n2, i2, k2, o2, a2)) in enumerate(zip(runs["plain"], runs["fused"])):
    assert n1 == n2, (n1, n2)
    rs = [rel(b, a) for a, b in zip(o1, o2) if a is not None and a.numel() > 0]
    flag = any(r > 1e-4 for r in rs)
    if flag or shown:
        print(i, n1, i1, k1, " ".join(f"{r:.2e}" for r in rs))
        if a1 is not None: print("      input diffs:", " ".join(f"{rel(b, a):.2e}" for a, b in zip(a1, a2)))
        shown += 1
        if shown > 3: break

print("== forward calls (fused run has conv_fwd_stats + bn_finalize where the plain one has conv_fwd + bn_stats)")
def key(e): return e[0].replace("conv_fwd_stats", "conv_fwd").replace("bn_finalize", "bn_stats")
shown = 0
for i, (e1, e2) in enumerate(zip(fwd["plain"], fwd["fused"])):
    assert key(e1) == key(e2), (e1[0], e2[0])
    o1, o2 = e1[3], e2[3]
    if e2[0] == "conv_fwd_stats": o2 = o2[:1]
    rs = [rel(b, a) for a, b in zip(o1, o2) if a is not None and b is not None and a.numel() > 0 and a.shape == b.shape]
    if any(r > 1e-5 for r in rs) or (shown and shown < 6):
        print(i, e1[0], e2[0], e1[1], " ".join(f"{r:.2e}" for r in rs)); shown += 1
