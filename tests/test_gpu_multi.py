"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box): the data-parallel path over NCCL with the CUDA library underneath --
synchronised BatchNorm on a batch sharded over two ranks against the unsharded batch on one GPU, the same check the
gloo test makes on CPU with the stage oracle (tests/test_distributed_cpu.py)."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out, layers):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import copy
        from fusion_gcn_b200 import graph as G, modules as M
        from fusion_gcn_b200.distributed import GradientAllReducer, SyncBatchNorm, shard_batch
        from helpers import ZERO_GRAD, rel_err, stat_err
        from oracle import agcn_oracle as O
        # full width, the oracle's loud initialisation; `layers` units
        shape, ncls, n_global = (2, 40, 25, 3), 12, 8
        torch.manual_seed(5)                                       # same model and data on every rank
        graph = G.SkeletonGraph(G.NTU_EDGES, center_joint=G.NTU_CENTER)
        state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=64, num_layers=layers, seed=3, loud=True)
        model = M.Model(shape, ncls, graph, start_feature_size=64, num_layers=layers)
        model.load_state_dict(state, strict=True)
        model = model.to(dev).train()
        x = torch.randn(n_global, *shape, device=dev)
        w = torch.randn(n_global, ncls, device=dev)
        full = copy.deepcopy(model)                                # the unsharded batch on one GPU: what the reference computes
        sharded = M.set_sync_batchnorm(copy.deepcopy(model), SyncBatchNorm())
        lo, hi = shard_batch(n_global, rank, world)
        reducer = GradientAllReducer(sharded.parameters())
        y = sharded(x[lo:hi])
        (y * w[lo:hi]).sum().backward()
        reducer()                                                  # mean over the ranks of the per-rank sums = full-batch gradient / world
        y_full = full(x)
        (y_full * w).sum().backward()
        torch.cuda.synchronize()
        err_y = rel_err(y, y_full[lo:hi])
        scale = max(float(q.grad.abs().max()) for q in full.parameters())
        # Two metrics per tensor.  The sharded and the unsharded run sum their statistics in different orders, so a ReLU input within
        # rounding of zero may take the other bracket (~2e7 ReLU inputs here); ONE flipped element moves a gradient tensor by up to
        # O(1e-2) of its maximum ("ReLU ties") but by almost nothing in the L2 norm, while a wrong sum or row count in
        # the synchronised backward moves every upstream gradient by O(1) in both.
        worst, worst_l2, names = 0.0, 0.0, []
        for (k, p), q in zip(sharded.named_parameters(), full.parameters()):
            g = p.grad.double() * world
            if ZERO_GRAD.search(k):
                assert float((g - q.grad).abs().max()) <= 1e-5 * scale, k
                continue
            e_max = rel_err(g, q.grad)
            e_l2 = float((g - q.grad.double()).norm() / q.grad.double().norm().clamp_min(1e-30))
            names.append((e_max, e_l2, k))
            worst, worst_l2 = max(worst, e_max), max(worst_l2, e_l2)
        # the same shards WITHOUT synchronisation: must be far outside the bound (the test has power)
        plain = copy.deepcopy(model)
        (plain(x[lo:hi]) * w[lo:hi]).sum().backward()
        g_plain = torch.cat([p.grad.reshape(-1) for p in plain.parameters()])
        dist.all_reduce(g_plain)
        g_full = torch.cat([q.grad.reshape(-1) for q in full.parameters()])
        unsync_l2 = float((g_plain.double() - g_full.double()).norm() / g_full.double().norm())
        stats = max(stat_err(a, b) for (k, a), b in zip(sharded.state_dict().items(), full.state_dict().values()) if "running" in k)
        # conditioning of the model itself: the SAME unsharded batch on the FFMA kernels (another summation order, every product exact in
        # fp32) -- how far two correct fp32 evaluations of these gradients are apart
        calib = M.set_precision(copy.deepcopy(model), "fp32_ffma")
        (calib(x) * w).sum().backward()
        noise_l2 = max(float((c.grad.double() - q.grad.double()).norm() / q.grad.double().norm().clamp_min(1e-30))
                       for (k, c), q in zip(calib.named_parameters(), full.parameters()) if not ZERO_GRAD.search(k))
        if rank == 0:
            names.sort(reverse=True)
            out.put((float(err_y), float(worst), float(worst_l2), float(stats), unsync_l2, noise_l2, names[:4]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
@pytest.mark.parametrize("layers", [5, 10])
def test_two_gpu_sync_batchnorm_matches_the_unsharded_batch(layers):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, layers)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=560)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    err_y, worst, worst_l2, stats, unsync_l2, noise_l2, names = out.get(timeout=5)
    print(f"[sync-bn over NCCL, {layers} units] y {err_y:.2e}, gradients: worst max-norm {worst:.2e}, worst L2 {worst_l2:.2e}, running statistics "
          f"{stats:.2e}; unsynchronised shards L2 {unsync_l2:.2e}; two fp32 evaluations of the unsharded batch L2 {noise_l2:.2e}; "
          f"worst tensors {[(f'{a:.1e}', f'{b:.1e}', k) for a, b, k in names]}")
    assert err_y <= 1e-5 and stats <= 1e-5, (err_y, stats)
    # the gradients of the sharded, synchronised run against the unsharded batch: as close as two correct fp32 evaluations of the
    # unsharded batch are to each other (loud-init stacks on 8 sequences amplify forward rounding differences of 4e-7 and ReLU ties),
    # and orders of magnitude closer than the unsynchronised shards
    assert worst_l2 <= max(1e-4, 4 * noise_l2), (worst_l2, noise_l2, names)
    assert unsync_l2 >= 20 * worst_l2, (unsync_l2, worst_l2)
