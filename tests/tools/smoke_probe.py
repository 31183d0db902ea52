"""Smoke-model gradient errors per tensor (debug aid): python tests/tools/smoke_probe.py [start] [seed]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from fusion_gcn_b200 import graph as G, modules as M
from oracle import agcn_oracle as O

start = int(sys.argv[1]) if len(sys.argv) > 1 else 16
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 1
shape, n, ncls = ((2, 16, 25, 3), 1, 60) if seed == 34 else ((2, 32, 25, 3), 2, 60)
graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, seed=seed, loud=True)
gen = torch.Generator().manual_seed(2)
x = torch.randn(n, *shape, generator=gen)
w = torch.randn(n, ncls, generator=gen)
model = M.Model(shape, ncls, graph, start_feature_size=start)
model.load_state_dict(state, strict=True)
model.cuda().train()
y = model(x.cuda())
(y * w.cuda()).sum().backward()
torch.cuda.synchronize()
p = O.as_leaves(state, torch.float64)
y_ref = O.model_forward(x.double(), p, 3, True, start=start)
(y_ref * w.double()).sum().backward()
errs = {k: O.rel_err(q.grad, p[k].grad) for k, q in model.named_parameters() if k.endswith("weight") or "adj_b" in k}
print("y err", O.rel_err(y, y_ref), " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("AGCN_")))
for k, v in sorted(errs.items(), key=lambda kv: -kv[1])[:4]:
    print(f"   {v:.3e}  {k}  maxabs ref {float(p[k].grad.abs().max()):.3e}")
