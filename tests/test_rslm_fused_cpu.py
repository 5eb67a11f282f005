"""The single-launch random-sample LM initialiser (epnp_rslm_f32, thread <-> hypothesis) against the unfused path
(gathered mini-problems -> epnp_lm_solve_f32 -> stacked evaluate_pnp), both running the real kernel source under the
CPU SIMT emulator.  Same hypotheses in, so every hypothesis' refined pose and full-set cost must agree up to fp32
summation order, and the selected pose per object must be the same."""
import numpy as np
import pytest
import torch

import simt_native
from conftest import err_vs
from epropnp.camera import PerspectiveCamera
from epropnp.common import evaluate_pnp
from epropnp.cost_fun import AdaptiveHuberPnPCost
from epropnp.levenberg_marquardt import RSLMSolver
from epropnp_b200 import native
from epropnp_b200.synth import make_problem


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


@pytest.mark.parametrize("dof,n,fast,bounded,N", [(6, 8, False, False, 64), (6, 12, True, True, 51), (4, 16, False, True, 64),
                                                (4, 5, True, False, 33)])
def test_every_hypothesis_matches_the_unfused_path(dev, dof, n, fast, bounded, N):
    B, P = 3, 9
    pc = {k: v.to(dev) for k, v in make_problem(B, N, seed=40 + n, dof=dof, outlier_frac=0.1).items()}
    x3d, x2d, w2d = pc["x3d"], pc["x2d"], pc["w2d"]
    lb, ub = ((x2d.amin(1) - 5.0, x2d.amax(1) - 20.0) if bounded else (None, None))     # ub cuts into the points
    delta = native.adaptive_delta(x2d, w2d, 0.5)
    prob = native.Problem(x3d, x2d, w2d, pc["cam_mats"], lb, ub, delta)
    g = torch.Generator().manual_seed(n)
    inds = torch.stack([torch.stack([torch.randperm(N, generator=g)[:n] for _ in range(B)]) for _ in range(P)]).to(dev)
    D = 7 if dof == 6 else 4
    start = pc["pose_init"][None].repeat(P, 1, 1) + 0.05 * torch.randn(P, B, D, generator=g).to(dev)
    if dof == 6:
        start[..., 3:] = torch.nn.functional.normalize(start[..., 3:], dim=-1)
    params = native.default_params(dof, lm_iter=3, fast_mode=int(fast))
    fused = native.rslm(prob, inds, start, params, want_all=True)
    # unfused: gather the mini-problems, solve them as P*B objects, score on the full sets
    rows = torch.arange(B, device=dev)[None, :, None]
    mini = native.Problem(x3d[rows, inds].reshape(P * B, n, 3), x2d[rows, inds].reshape(P * B, n, 2),
                          w2d[rows, inds].reshape(P * B, n, 2), pc["cam_mats"].repeat(P, 1, 1),
                          None if lb is None else lb.repeat(P, 1), None if ub is None else ub.repeat(P, 1),
                          delta.repeat(P))
    pose = native.lm_solve(mini, start.reshape(P * B, D), params)["pose_opt"].reshape(P, B, D)
    cost = native.evaluate_cost(prob, pose, dof, params.z_min)
    close = (fused["pose_all"] - pose).abs().amax(-1) < 1e-4 * pose.abs().amax()
    assert close.float().mean() >= 0.9                    # the rest: an accept / reject decided the other way round
    assert err_vs(fused["cost_all"][close].cpu().numpy(), cost[close].cpu().numpy()) < 1e-4
    # the selection is the argmin of the fused costs, ties and all
    best = fused["cost_all"].argmin(dim=0)
    assert torch.equal(fused["cost"], fused["cost_all"][best, torch.arange(B, device=dev)])
    assert torch.equal(fused["pose"], fused["pose_all"][best, torch.arange(B, device=dev)])


def test_solver_class_returns_the_cheapest_hypothesis(dev):
    B, N = 4, 64
    pc = {k: v.to(dev) for k, v in make_problem(B, N, seed=50).items()}
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"])
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(pc["x2d"], pc["w2d"])
    solver = RSLMSolver(dof=6, num_points=8, num_proposals=16, num_iter=5)
    torch.manual_seed(123)
    p1, n1, c1 = solver.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
    assert n1 is None and p1.shape == (B, 7) and c1.shape == (B,)
    full = evaluate_pnp(pc["x3d"], pc["x2d"], pc["w2d"], p1, camera, cost_fun, out_cost=True)[1]
    assert torch.allclose(full, c1, rtol=1e-5, atol=1e-5)
    torch.manual_seed(123)                                 # same draws -> same answer, bit for bit
    p2, _, c2 = solver.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
    assert torch.equal(p1, p2) and torch.equal(c1, c2)


def test_nan_hypothesis_wins_like_torch_min(dev):
    """cost.min(dim=0) propagates NaN (levenberg_marquardt.py:350); the fused selection does the same."""
    B, N, P, n = 2, 32, 4, 6
    pc = {k: v.to(dev) for k, v in make_problem(B, N, seed=51).items()}
    x2d = pc["x2d"].clone()
    x2d[1, 3, 0] = float("nan")                           # object 1: every full-set cost is NaN
    prob = native.Problem(pc["x3d"], x2d, pc["w2d"], pc["cam_mats"], None, None, torch.ones(B, device=dev))
    g = torch.Generator().manual_seed(1)
    inds = torch.stack([torch.stack([torch.randperm(N, generator=g)[:n] for _ in range(B)]) for _ in range(P)]).to(dev)
    out = native.rslm(prob, inds, pc["pose_init"][None].repeat(P, 1, 1), native.default_params(6, lm_iter=2), want_all=True)
    assert torch.isfinite(out["cost"][0]) and torch.isnan(out["cost"][1])
    assert np.isnan(out["cost_all"][:, 1].cpu().numpy()).all()
