"""Fused solve + in-kernel gather (epnp_lm_amis_fused_push_f32: the AMIS kernel's push epilogue) under the CPU SIMT emulator: two
"ranks" solve the two halves of a batch, each pushing its rows into the other's full-batch buffers (plain host memory
standing in for IPC-mapped peer memory).  Every rank's assembled buffer must equal the single-"GPU" run bit for bit,
and rows the rank does not own must not be touched by its own launch."""
import pytest
import torch

import simt_native
from epropnp_b200 import native
from epropnp_b200.synth import make_problem


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


def shard(pc, lo, hi):
    return {k: v[lo:hi].contiguous() for k, v in pc.items()}


@pytest.mark.parametrize("dof,N,M,I", [(6, 24, 16, 2), (4, 21, 12, 3), (6, 9, 10, 2)])
def test_pushed_rows_assemble_the_single_gpu_batch(dev, dof, N, M, I):
    B, world = 10, 2
    D = 7 if dof == 6 else 4
    pc = make_problem(B, N, seed=21, dof=dof)
    p = native.default_params(dof, lm_iter=3, mc_samples=M, mc_iter=I)
    delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    whole = native.lm_amis_fused(native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta),
                                 pc["pose_init"], p, seed=5, obj_offset=0, want_cov=False, want_cost_init=False)
    poison = float("nan")
    full_logw = [torch.full((B, M), poison) for _ in range(world)]
    full_pose = [torch.full((B, D), poison) for _ in range(world)]
    per = B // world
    for r in range(world):
        lo, hi = r * per, (r + 1) * per
        sh = shard(pc, lo, hi)
        prob = native.Problem(sh["x3d"], sh["x2d"], sh["w2d"], sh["cam_mats"], None, None, delta[lo:hi].contiguous())
        peers = [q for q in range(world) if q != r]
        out = native.lm_amis_fused_push(prob, sh["pose_init"], p, full_pose[r][lo:hi], full_logw[r][lo:hi],
                                        [full_logw[q] for q in peers], [full_pose[q] for q in peers],
                                        seed=5, obj_offset=lo)
        assert torch.equal(out["pose_samples"], whole["pose_samples"][lo:hi])
        # after rank r's launch: its rows are present everywhere, nobody else's rows were written by it
        for q in range(world):
            assert torch.equal(full_logw[q][lo:hi], whole["logw"][lo:hi])
            assert torch.equal(full_pose[q][lo:hi], whole["pose_opt"][lo:hi])
        if r == 0:
            assert torch.isnan(full_logw[1][per:]).all() and torch.isnan(full_pose[0][per:]).all()
    for q in range(world):
        assert torch.equal(full_logw[q], whole["logw"]) and torch.equal(full_pose[q], whole["pose_opt"])


def test_push_without_peers_is_the_plain_fused_call(dev):
    pc = make_problem(4, 16, seed=3)
    p = native.default_params(6, lm_iter=2, mc_samples=8, mc_iter=2)
    delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta)
    ref = native.lm_amis_fused(prob, pc["pose_init"], p, seed=9, want_cost=True, want_cov=True, want_cost_init=False)
    out = native.lm_amis_fused_push(prob, pc["pose_init"], p, torch.empty(4, 7), torch.empty(4, 8), [], [], seed=9,
                                    want_cost=True, want_cov=True)
    for k in ("pose_opt", "logw", "pose_samples", "cost", "pose_cov"):
        assert torch.equal(out[k], ref[k]), k


def test_push_argument_checks(dev):
    pc = make_problem(4, 16, seed=3)
    p = native.default_params(6, lm_iter=2, mc_samples=8, mc_iter=2)
    delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta)
    with pytest.raises(ValueError):                                    # peer buffer too short for obj_offset + B
        native.lm_amis_fused_push(prob, pc["pose_init"], p, torch.empty(4, 7), torch.empty(4, 8),
                                  [torch.empty(5, 8)], [torch.empty(5, 7)], obj_offset=2)
    with pytest.raises(ValueError):                                    # more than 8 peers
        native.lm_amis_fused_push(prob, pc["pose_init"], p, torch.empty(4, 7), torch.empty(4, 8),
                                  [torch.empty(4, 8)] * 9, [torch.empty(4, 7)] * 9)


# ------------------------------------------------------------------------------------------------
# sharded.PushGather's protocol (ring slots, row ranges, gates, rendezvous) in two REAL processes over gloo: the kernel
# runs under the SIMT emulator and the peers' buffers are shared-memory host tensors standing in for CUDA-IPC mappings.
# Streams and events are inert on the host (everything executes in program order), so this checks the bookkeeping --
# which slot, which rows, which gate -- not the asynchronous hazards; those need tests/test_push_gather_gpu.py.
class _HostStream:
    def wait_event(self, e):
        pass

    def wait_stream(self, s):
        pass


class _HostEvent:
    def record(self, stream=None):
        pass


def _host_class(base):
    import contextlib
    from torch.multiprocessing.reductions import reduce_storage

    class HostGather(base):
        def _new_stream(self):
            return _HostStream()

        def _new_event(self):
            return _HostEvent()

        def _current_stream(self):
            return _HostStream()

        def _on_stream(self, stream):
            return contextlib.nullcontext()

        def _device_synchronize(self):
            pass

        def _record_stream(self, tensor, stream):
            pass

        # PushGather maps its peers through raw CUDA IPC on the GPU; on the host shared-memory tensors stand in.
        # reduce_tensor() of a HOST tensor embeds the storage object (only ForkingPickler shares it); export the
        # shared-memory file of the storage itself so that all_gather_object's plain pickle carries a real handle
        def _export_raw(self, t):
            t.share_memory_()
            fn, args = reduce_storage(t.untyped_storage())
            return (fn, args, tuple(t.shape), t.dtype)

        def _import_raw(self, handle):
            fn, args, shape, dtype = handle
            return torch.empty(0, dtype=dtype).set_(fn(*args), 0, shape)

    return HostGather


def _host_push_gather_class():
    from epropnp_b200.sharded import PushGather
    return _host_class(PushGather)


def _push_worker(rank, world, port, q):
    import contextlib
    import os
    import torch.distributed as dist
    from epropnp_b200 import capi
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.multiprocessing.set_sharing_strategy("file_system")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        capi._lib = simt_native.handle(())                       # what simt_native.install does, without pytest
        native._need_cuda = lambda t, what: None
        native.stream_ptr = lambda device=None: None
        torch.cuda.device = lambda device=None: contextlib.nullcontext()
        per, N, M, depth = 3, 16, 8, 3
        num_obj = per * world
        p = native.default_params(6, lm_iter=2, mc_samples=M, mc_iter=2)
        pg = _host_push_gather_class()(num_obj, M, 7, "cpu", depth=depth, valid_for=2, copy_out=bool(rank))  # both forms
        lo, hi = rank * per, (rank + 1) * per
        ok, pending, expected = True, None, None
        for step in range(2 * depth + 1):                           # every slot reused twice
            pc = make_problem(num_obj, N, seed=70 + step)
            delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
            whole = native.lm_amis_fused(native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta),
                                         pc["pose_init"], p, seed=step, want_cov=False, want_cost_init=False)
            mine = native.Problem(pc["x3d"][lo:hi], pc["x2d"][lo:hi], pc["w2d"][lo:hi], pc["cam_mats"][lo:hi], None, None,
                                  delta[lo:hi])
            out, nxt = pg.solve(mine, pc["pose_init"][lo:hi], p, seed=step)
            ok = ok and torch.equal(out["pose_samples"], whole["pose_samples"][lo:hi])
            if pending is not None:                                  # read batch t-1 after batch t was started
                got = pending.wait()
                ok = ok and torch.equal(got["logw"], expected["logw"]) and torch.equal(got["pose_opt"], expected["pose_opt"])
            pending, expected = nxt, dict(logw=whole["logw"].clone(), pose_opt=whole["pose_opt"].clone())
        got = pending.wait()
        ok = ok and torch.equal(got["logw"], expected["logw"]) and torch.equal(got["pose_opt"], expected["pose_opt"])
        dist.barrier()
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_push_gather_protocol_two_processes_gloo():
    import socket
    import torch.multiprocessing as mp
    simt_native.build_emulated(())                                   # build once, before the workers race for it
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_push_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    results = [q.get(timeout=240) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(results) == [(r, True) for r in range(world)]

def test_push_and_epilogue_do_not_depend_on_thread_schedule(dev):
    """Missing-barrier probe (as tests/test_simt_emul_cpu.py does for the validated kernels): the fibers of a CTA run in
    descending and randomly permuted order; the pushed rows and the epilogue outputs must stay bit-identical."""
    lib = simt_native.handle(())
    pc = make_problem(5, 37, seed=12)
    M = 16
    p = native.default_params(6, lm_iter=3, mc_samples=M, mc_iter=2)
    delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], None, None, delta)

    def run():
        peers_lw, peers_ps = [torch.zeros(7, M) for _ in range(2)], [torch.zeros(7, 7) for _ in range(2)]
        out = native.lm_amis_fused_push(prob, pc["pose_init"], p, torch.empty(5, 7), torch.empty(5, M), peers_lw, peers_ps,
                                        seed=3, obj_offset=2)
        ep = native.mc_epilogue(out["logw"], out["pose_samples"], out["pose_opt"], cost_target=torch.ones(5),
                                want_lse=True, want_loss=True, want_weights=True, want_score=True)
        g = native.mc_lse_backward(out["logw"], ep["lse"], torch.ones(5))
        return peers_lw + peers_ps + [out["logw"], out["pose_opt"], ep["lse"], ep["loss"], ep["weights"], ep["score_te"], g]
    try:
        lib.simt_set_schedule(0, 1)
        base = run()
        for mode, seed in ((1, 1), (2, 11), (2, 12)):
            lib.simt_set_schedule(mode, seed)
            for a, b in zip(base, run()):
                assert torch.equal(a, b), (mode, seed)
    finally:
        lib.simt_set_schedule(0, 1)
