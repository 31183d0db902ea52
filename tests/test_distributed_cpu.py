"""World-size-2 data parallelism on CPU (gloo): batch sharding + bucketed gradient all-reduce.

Each rank runs the host-side composition on its batch shard (on top of the TEST-ONLY torch stage backend, because there
is no GPU here), all-reduces the gradients with ``GradientAllReducer`` and compares them with the mean of the per-shard
gradients of the CPU oracle.  BatchNorm statistics are per replica, exactly like DistributedDataParallel.
"""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = 5          # l0..l4 (includes the first stride-2 / channel-doubling unit); deeper loud-init stacks on tiny shards are ill-conditioned


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_global, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fusion_gcn_b200.functional as FN
        from fusion_gcn_b200 import graph as G, modules as M
        from fusion_gcn_b200.distributed import GradientAllReducer, shard_batch
        from oracle import agcn_oracle as O, stages
        FN.K = stages                                   # TEST ONLY: torch stage oracle instead of the CUDA library
        shape, ncls, start = (1, 12, 20, 3), 7, 8
        graph = G.SkeletonGraph(G.UTD_EDGES, center_joint=G.UTD_CENTER)
        state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, num_layers=LAYERS, seed=3, loud=True)
        gen = torch.Generator().manual_seed(11)
        x = torch.randn(n_global, *shape, generator=gen)
        w = torch.randn(n_global, ncls, generator=gen)
        lo, hi = shard_batch(n_global, rank, world)
        model = M.Model(shape, ncls, graph, start_feature_size=start, num_layers=LAYERS)
        model.load_state_dict(state, strict=True)
        model.train()
        reducer = GradientAllReducer(model.parameters(), bucket_bytes=4096)       # small buckets: several per step
        (model(x[lo:hi]) * w[lo:hi]).sum().backward()
        reducer()
        # expected: mean over ranks of the oracle's gradient on each rank's shard (per-replica BN statistics)
        def mean_of_shard_grads(dtype):
            total = None
            for r in range(world):
                a, b = shard_batch(n_global, r, world)
                p = O.as_leaves(state, dtype)
                (O.model_forward(x[a:b].to(dtype), p, shape[3], True, start=start, num_layers=LAYERS) * w[a:b].to(dtype)).sum().backward()
                g = {k: v.grad for k, v in p.items() if v.requires_grad and v.grad is not None}
                total = g if total is None else {k: total[k] + g[k] for k in g}
            return {k: v / world for k, v in total.items()}

        expect = mean_of_shard_grads(torch.float64)
        expect32 = mean_of_shard_grads(torch.float32)      # the reference arithmetic's own fp32 noise (tiny BN batches here)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import check_grads                # SURVEY D8 metric (zero-gradient families on an absolute bound)
        _, worst = check_grads({k: prm.grad for k, prm in model.named_parameters()},
                               {k: expect[k] for k, _ in model.named_parameters()}, 1e-4, f"rank {rank}",
                               ref32={k: expect32[k] for k, _ in model.named_parameters()})
        # every rank must hold identical averaged gradients
        flat = torch.cat([prm.grad.reshape(-1) for prm in model.parameters()])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        same = all(torch.equal(gathered[0], t) for t in gathered)
        if rank == 0:
            out.put((worst, same, len(reducer.buckets)))
    finally:
        dist.destroy_process_group()


def test_shard_batch_partitions():
    from fusion_gcn_b200.distributed import shard_batch
    assert [shard_batch(64, r, 4) for r in range(4)] == [(0, 16), (16, 32), (32, 48), (48, 64)]
    with pytest.raises(ValueError):
        shard_batch(10, 0, 4)


@pytest.mark.timeout(300)
def test_two_rank_gradient_allreduce_matches_mean_of_shard_gradients():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=280)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    worst, same, nbuckets = out.get(timeout=5)
    assert same, "ranks disagree on the all-reduced gradients"
    assert worst <= 5e-2, worst          # per-tensor bounds are asserted inside the workers (check_grads)
    assert nbuckets > 1


def _sync_worker(rank, world, port, n_global, out):
    """Synchronised BatchNorm: the batch sharded over the ranks must normalise exactly like the unsharded batch."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fusion_gcn_b200.functional as FN
        from fusion_gcn_b200 import graph as G, modules as M
        from fusion_gcn_b200.distributed import GradientAllReducer, SyncBatchNorm, shard_batch
        from oracle import agcn_oracle as O, stages
        FN.K = stages                                   # TEST ONLY: torch stage oracle instead of the CUDA library
        shape, ncls, start = (2, 12, 20, 3), 7, 8
        graph = G.SkeletonGraph(G.UTD_EDGES, center_joint=G.UTD_CENTER)
        state = O.init_state(G.adjacency_from_graph(graph), shape, ncls, start=start, num_layers=LAYERS, seed=3, loud=True)
        gen = torch.Generator().manual_seed(11)
        x = torch.randn(n_global, *shape, generator=gen)
        w = torch.randn(n_global, ncls, generator=gen)
        lo, hi = shard_batch(n_global, rank, world)
        model = M.Model(shape, ncls, graph, start_feature_size=start, num_layers=LAYERS)
        model.load_state_dict(state, strict=True)
        M.set_sync_batchnorm(model, SyncBatchNorm())
        M.set_recompute(model, rank == 1)               # the policies compose: one rank recomputes theta / phi and z, the other keeps them
        assert model._agcn_sync is not None and model.l0.gcn1._agcn_sync is model._agcn_sync and model._agcn_sync.world == world
        model.train()
        reducer = GradientAllReducer(model.parameters(), bucket_bytes=4096)
        y = model(x[lo:hi])
        (y * w[lo:hi]).sum().backward()
        reducer()

        def full_batch(dtype):
            p = O.as_leaves(state, dtype)
            yf = O.model_forward(x.to(dtype), p, shape[3], True, start=start, num_layers=LAYERS)
            (yf * w.to(dtype)).sum().backward()
            return yf.detach(), {k: v.grad / world for k, v in p.items() if v.requires_grad and v.grad is not None}

        y64, expect = full_batch(torch.float64)
        _, expect32 = full_batch(torch.float32)
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from helpers import check_grads, rel_err
        err_y = rel_err(y, y64[lo:hi])
        assert err_y <= 2e-5, err_y
        _, worst = check_grads({k: prm.grad for k, prm in model.named_parameters()},
                               {k: expect[k] for k, _ in model.named_parameters()}, 1e-4, f"sync rank {rank}",
                               ref32={k: expect32[k] for k, _ in model.named_parameters()})
        # running statistics: those of the whole batch, identical on every rank
        stats = torch.cat([b.reshape(-1).float() for k, b in model.named_buffers() if "running" in k])
        gathered = [torch.empty_like(stats) for _ in range(world)]
        dist.all_gather(gathered, stats)
        same = all(torch.allclose(gathered[0], t, rtol=0, atol=1e-6) for t in gathered)
        # ... and NOT what per-replica statistics would have given
        plain = M.Model(shape, ncls, graph, start_feature_size=start, num_layers=LAYERS)
        plain.load_state_dict(state, strict=True)
        plain.train()
        differs = rel_err(plain(x[lo:hi]), y64[lo:hi]) > 1e-3
        if rank == 0:
            out.put((worst, same, differs, float(err_y)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_sync_batchnorm_matches_the_unsharded_batch():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, 4, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=280)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    worst, same, differs, err_y = out.get(timeout=5)
    assert same, "ranks disagree on the running statistics"
    assert differs, "per-replica statistics gave the same output: the test does not exercise the synchronisation"
    assert worst <= 5e-2 and err_y <= 2e-5, (worst, err_y)
