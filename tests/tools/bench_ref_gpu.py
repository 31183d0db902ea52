"""Same-hardware baseline: the oracle port of the reference (eager PyTorch: cuDNN conv, cuBLAS bmm, ATen softmax / BN)
on one B200, NTU shape, fwd+CE+bwd, train mode.  Reported beside the product numbers, never used by the product.

  python tests/tools/bench_ref_gpu.py [--batch 64] [--steps 5] [--tf32]     (run on the GPU box)
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--tf32", action="store_true", help="PyTorch's default on B200: TF32 cuDNN convs (bmm stays fp32)")
    args = ap.parse_args()
    from fusion_gcn_b200 import graph as G
    from oracle import agcn_oracle as O
    torch.backends.cudnn.allow_tf32 = args.tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.benchmark = True
    m, t, v, c, ncls = 2, 300, 25, 3, 60
    adj = G.adjacency_from_graph(G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER))
    state = O.init_state(adj, (m, t, v, c), ncls, seed=1)
    p = {k: (a.cuda() if torch.is_tensor(a) else a) for k, a in O.as_leaves(state).items()}
    for a in p.values():
        if a.is_floating_point() and a.is_leaf and not a.requires_grad:
            pass
    leaves = []
    for k in list(p):
        if state[k].is_floating_point() and not (k.endswith("running_mean") or k.endswith("running_var") or k.endswith("adj_a")):
            p[k] = p[k].detach().requires_grad_(True)
            leaves.append(p[k])
    gen = torch.Generator().manual_seed(1234)
    x = torch.randn(args.batch, m, t, v, c, generator=gen).cuda()
    y = torch.randint(ncls, (args.batch,), generator=gen).cuda()

    def step():
        for a in leaves:
            a.grad = None
        loss = torch.nn.functional.cross_entropy(O.model_forward(x, p, c, True), y)
        loss.backward()
        return loss

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"what": "oracle port, eager PyTorch on cuda:0 (same-hardware baseline)", "cudnn_tf32": args.tf32,
                      "batch": args.batch, "ms_per_step": round(ms, 3), "sequences_per_s": round(args.batch / ms * 1e3, 2),
                      "loss": round(float(loss), 5), "peak_mem_gb": round(torch.cuda.max_memory_allocated() / 2**30, 2)}))


if __name__ == "__main__":
    main()
