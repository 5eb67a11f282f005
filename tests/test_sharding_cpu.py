"""World-size-2 (and 3, ragged) gloo test of the batch sharding + gather logic -- no GPU needed."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from epropnp_b200.sharded import gather_objects, gather_results, gather_results_async, shard_range, shard_sizes


def test_shard_range_partitions():
    for n in (0, 1, 7, 8, 4096, 32768, 1001):
        for w in (1, 2, 3, 4, 8):
            edges = [shard_range(n, r, w) for r in range(w)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(w - 1))
            sizes = shard_sizes(n, w)
            assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, num_obj, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full_pose = torch.arange(num_obj * 7, dtype=torch.float32).reshape(num_obj, 7)
        full_logw = torch.arange(num_obj * 16, dtype=torch.float32).reshape(num_obj, 16) * 0.5
        b, e = shard_range(num_obj, rank, world)
        local = dict(pose_opt=full_pose[b:e].clone(), logw=full_logw[b:e].clone(), pose_cov=None, cost=None)
        got = gather_results(local, num_obj)
        ok = torch.equal(got["pose_opt"], full_pose) and torch.equal(got["logw"], full_logw) and "cost" not in got
        ok = ok and torch.equal(gather_objects(full_pose[b:e].clone(), num_obj), full_pose)
        pend = gather_results_async(local, num_obj)
        got2 = pend.wait()
        ok = ok and torch.equal(got2["pose_opt"], full_pose) and torch.equal(got2["logw"], full_logw)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,num_obj", [(2, 8), (2, 9), (3, 10)])
def test_gather_gloo(world, num_obj):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, num_obj, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]


def _loss_worker(rank, world, port, q):
    """Detection flavour of MonteCarloPoseLoss: the EMA norm factor is fed with the MEAN over ranks (mmdet reduce_mean,
    EPro-PnP-Det/epropnp_det/models/losses/monte_carlo_pose_loss.py:52-55), so every rank holds the same buffer."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "epro-pnp_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from epropnp.monte_carlo_pose_loss import MonteCarloPoseLoss
        g = torch.Generator().manual_seed(rank)
        logw, ct = torch.randn(32, 5, generator=g), torch.rand(5, generator=g)
        mod = MonteCarloPoseLoss(loss_weight=0.15, init_norm_factor=1.0, momentum=0.1, sync_norm_factor=True).train()
        loss = mod(logw, ct, torch.tensor(float(rank + 1)))
        expect_nf = 0.9 * 1.0 + 0.1 * (sum(range(1, world + 1)) / world)
        expect = (ct + torch.logsumexp(logw, dim=0)).mean() * (0.15 / expect_nf)
        ok = abs(mod.norm_factor.item() - expect_nf) < 1e-6 and abs(loss.item() - expect.item()) < 1e-5
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_loss_norm_factor_is_averaged_over_ranks_gloo():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_loss_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
