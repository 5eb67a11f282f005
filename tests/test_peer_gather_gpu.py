"""2-GPU test of the copy-engine gather (sharded.PeerGather): pulls from IPC-mapped peer buffers must reproduce the
NCCL all-gather bit for bit over several batches (ring-slot reuse included).  Needs two GPUs on one node; skipped
on a single-GPU box.  PeerGather is experimental and has not had its first hardware run yet, so the test is also
gated by the environment:
    gpurun --gpus 2 -- 'EPNP_TEST_PEER_GATHER=1 timeout 300 python -m pytest tests/test_peer_gather_gpu.py -q'
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from epropnp_b200.sharded import PeerGather, gather_results
        per, M = 64, 128
        num_obj = per * world
        g = torch.Generator(device=dev).manual_seed(100 + rank)
        pg, ok, pending, expected = None, True, None, None
        for step in range(7):                                   # > depth: every ring slot is reused
            local = dict(pose_opt=torch.randn(per, 7, device=dev, generator=g),
                         logw=torch.randn(per, M, device=dev, generator=g))
            want = gather_results(local, num_obj, keys=("pose_opt", "logw"))
            if pg is None:
                pg = PeerGather(local, num_obj, keys=("pose_opt", "logw"), depth=2)
            if pending is not None:                             # overlapped use: wait for batch t-1 after starting t
                nxt = pg.start(local)
                got = pending.wait()
                torch.cuda.synchronize()
                ok = ok and all(torch.equal(got[k], expected[k]) for k in expected)
                pending = nxt
            else:
                pending = pg.start(local)
            expected = want
        got = pending.wait()
        torch.cuda.synchronize()
        ok = ok and all(torch.equal(got[k], expected[k]) for k in expected)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_peer_gather_matches_nccl():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
