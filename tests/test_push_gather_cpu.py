"""Fused solve + in-kernel gather (epnp_lm_amis_fused_push_f32 / solve_push_kernel) under the CPU SIMT emulator: two
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
