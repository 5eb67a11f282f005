"""2-GPU test of the fused solve + gather (sharded.PushGather -> the AMIS kernel's push epilogue): every rank's full-batch buffers,
filled by all ranks' kernels through IPC-mapped peer memory, must equal the single-GPU solve of the whole batch bit for
bit, over more batches than the ring is deep (slot reuse) and in the overlapped use pattern (start batch t+1, then
read batch t).  Two GPUs on one node: one rank per GPU over NCCL, the stores cross NVLink
    gpurun --gpus 2 -- 'timeout 300 python -m pytest tests/test_push_gather_gpu.py -q'
On a single-GPU box the same two ranks share cuda:0 (rendezvous over gloo, NCCL refuses two ranks on one device): the
kernel of one PROCESS still stores into buffers of the other process mapped through CUDA IPC, ring, slots and gating are
the same -- only the wire is missing.
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


def _worker(rank, world, port, q, one_device):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dev = torch.device("cuda", 0 if one_device else rank)
    torch.cuda.set_device(dev)
    if one_device:
        dist.init_process_group("gloo", rank=rank, world_size=world)
    else:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from epropnp_b200 import native
        from epropnp_b200.sharded import PushGather
        from epropnp_b200.synth import make_problem
        per, N, M = 96, 64, 128
        num_obj = per * world
        p = native.default_params(6, lm_iter=5, mc_samples=M, mc_iter=4)
        pg = PushGather(num_obj, M, 7, dev, depth=3, valid_for=2)
        ok, pending, expected = True, None, None
        lo, hi = rank * per, (rank + 1) * per
        for step in range(8):                                   # > depth: every ring slot is reused
            pc = {k: v.to(dev) for k, v in make_problem(num_obj, N, seed=300 + step).items()}
            delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
            whole = native.lm_amis_fused(native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta),
                                         pc["pose_init"], p, seed=40 + step, want_cov=False, want_cost_init=False)
            mine = native.Problem(pc["x3d"][lo:hi], pc["x2d"][lo:hi], pc["w2d"][lo:hi], pc["cam_mats"][lo:hi], None, None,
                                  delta[lo:hi])
            out, nxt = pg.solve(mine, pc["pose_init"][lo:hi], p, seed=40 + step)
            ok = ok and torch.equal(out["pose_samples"], whole["pose_samples"][lo:hi])
            if pending is not None:                             # overlapped use: read batch t-1 after starting batch t
                got = pending.wait()
                torch.cuda.synchronize()
                ok = ok and torch.equal(got["logw"], expected["logw"]) and torch.equal(got["pose_opt"], expected["pose_opt"])
            pending, expected = nxt, dict(logw=whole["logw"].clone(), pose_opt=whole["pose_opt"].clone())
        got = pending.wait()
        torch.cuda.synchronize()
        ok = ok and torch.equal(got["logw"], expected["logw"]) and torch.equal(got["pose_opt"], expected["pose_opt"])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_push_gather_assembles_the_single_gpu_batch():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    world = 2
    one_device = torch.cuda.device_count() < 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, one_device)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]
