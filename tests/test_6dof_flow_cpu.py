"""The 6DoF repository's call sequences (EPro-PnP-6DoF/lib/train.py:157-193, lib/test.py:196-221) on the drop-in
package with the kernels under the CPU SIMT emulator: sub-sampled dense correspondences (advanced indexing, expanded
camera matrix, per-object tensor bounds), training forward with pose_init / force_init_solve / with_pose_opt_plus and
the three losses back-propagated, then the test-time Gauss-Newton path and the visualisation Monte-Carlo pass."""
import math

import numpy as np
import pytest
import torch

import simt_native
from epropnp.camera import PerspectiveCamera
from epropnp.cost_fun import AdaptiveHuberPnPCost
from epropnp.epropnp import EProPnP6DoF
from epropnp.levenberg_marquardt import LMSolver, RSLMSolver
from epropnp_b200.synth import make_problem


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


def _dense_batch(bs, res, seed):
    pc = make_problem(bs, res * res, seed=seed, grid2d=True)
    rng = np.random.RandomState(seed)
    keep = torch.from_numpy(np.stack([rng.choice(res * res, size=res * res // 4, replace=False) for _ in range(bs)]))
    rows = torch.arange(bs)[:, None]
    return pc, pc["x3d"][rows, keep], pc["x2d"][rows, keep], keep, rows


def test_training_step(dev):
    bs, res = 3, 16
    pc, x3d0, x2d, keep, rows = _dense_batch(bs, res, seed=31)
    x3d = x3d0.clone().requires_grad_(True)
    w_logit = (0.3 * torch.randn(bs, res * res, 2, generator=torch.Generator().manual_seed(1))).requires_grad_(True)
    scale = torch.full((bs, 2), 40.0, requires_grad=True)
    w2d = w_logit[rows, keep]
    w2d = (w2d - w2d.mean(dim=1, keepdim=True) - math.log(w2d.size(1))).exp() * scale[:, None, :]   # train.py:164
    lo, hi = x2d.amin(dim=1), x2d.amax(dim=1)
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"][0][None].expand(bs, -1, -1), z_min=0.01,
                               lb=lo - 30.0, ub=hi + 30.0)
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.1)
    cost_fun.set_param(x2d, w2d)
    assert cost_fun.delta.requires_grad                      # delta depends on w2d: gradients flow through it
    epropnp = EProPnP6DoF(mc_samples=128, num_iter=4,
                          solver=LMSolver(dof=6, num_iter=5,
                                          init_solver=RSLMSolver(dof=6, num_points=8, num_proposals=8, num_iter=3)))
    gt = pc["pose_gt"]
    torch.manual_seed(7)
    _, _, pose_opt_plus, _, logw, cost_tgt = epropnp.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=gt, force_init_solve=True, with_pose_opt_plus=True)
    loss_mc = (cost_tgt + torch.logsumexp(logw, dim=0)).mean() / 10.0                # monte_carlo_pose_loss.py:27-33
    loss_t = (pose_opt_plus[:, :3] - gt[:, :3]).norm(dim=-1)
    loss_t = torch.where(loss_t < 0.05, 0.5 * loss_t.square() / 0.05, loss_t - 0.025).mean()
    dot = (pose_opt_plus[:, None, 3:] @ gt[:, 3:, None]).squeeze(-1).squeeze(-1)
    loss_r = ((1 - dot.square()) * 2).mean()
    (0.02 * loss_mc + 0.1 * loss_t + 0.1 * loss_r).backward()
    for t in (x3d, w_logit, scale):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    unused = torch.ones(bs, res * res, dtype=torch.bool)
    unused[rows, keep] = False
    assert w_logit.grad[unused].abs().sum() == 0             # only the sampled correspondences receive gradient


def test_inference_step(dev):
    bs, res = 2, 16
    pc = make_problem(bs, res * res, seed=32, grid2d=True)
    x3d, x2d, w2d = pc["x3d"], pc["x2d"], pc["w2d"]
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"][0][None].expand(bs, -1, -1), z_min=0.01)
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.1)
    epropnp = EProPnP6DoF(mc_samples=128, num_iter=4, solver=LMSolver(dof=6, num_iter=3))
    with torch.no_grad():
        cost_fun.set_param(x2d, w2d)
        pose_opt = epropnp(x3d, x2d, w2d, camera, cost_fun, pose_init=pc["pose_init"], fast_mode=True)[0]
        _, _, _, samples, logw, _ = epropnp.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_opt,
                                                               force_init_solve=False, fast_mode=True)
    gt = pc["pose_gt"]
    assert ((pose_opt[:, :3] - gt[:, :3]).norm(dim=-1) < 0.05).all()
    assert (1 - (pose_opt[:, 3:] * gt[:, 3:]).sum(-1).abs() < 1e-3).all()
    assert samples[:, :1].shape == (128, 1, 7) and logw[:, :1].shape == (128, 1) and torch.isfinite(logw).all()


def test_demo_training_loop_runs(dev):
    """demo/fit_identity.py (the reference notebook's experiment) for a few optimiser steps on the emulated kernels:
    RSLM initialisation in every forward, fused LM + AMIS, native backward, Adam.  (That the loss falls is the GPU
    test's business -- tests/test_demo_gpu.py runs 160 steps.)"""
    import os
    import sys
    from conftest import ROOT
    sys.path.insert(0, os.path.join(ROOT, "demo"))
    import fit_identity
    out = fit_identity.run(steps=3, batch_size=4, verbose=False, device=dev, test_size=4)
    assert out["finite"] and out["steps"] == 3
    assert all(math.isfinite(out[k]) for k in ("loss_mc_first", "loss_mc_last", "test_t_err_before", "test_t_err_after"))


def test_demo_training_loop_with_native_backward_and_epilogue(dev, monkeypatch):
    """The same loop with the native pose_opt_plus backward and the native Monte-Carlo loss epilogue (through the package's
    MonteCarloPoseLoss) -- the GPU defaults -- against the torch composites the CPU harness normally runs.  Gradients and
    losses must stay finite and agree."""
    import os
    import sys
    from conftest import ROOT
    from epropnp import monte_carlo_pose_loss as mcl
    sys.path.insert(0, os.path.join(ROOT, "demo"))
    import fit_identity
    base = fit_identity.run(steps=2, batch_size=4, verbose=False, device=dev, test_size=4, seed=3)
    for k in ("EPNP_NATIVE_GN_STEP", "EPNP_NATIVE_MC_EPILOGUE"):
        monkeypatch.setenv(k, "1")
    monkeypatch.setattr(mcl, "_use_native", lambda t: True)
    monkeypatch.setattr(fit_identity, "MonteCarloPoseLoss", lambda: mcl.MonteCarloPoseLoss(momentum=0.1))
    out = fit_identity.run(steps=2, batch_size=4, verbose=False, device=dev, test_size=4, seed=3)
    assert out["finite"] and out["steps"] == 2
    assert math.isfinite(out["loss_mc_first"]) and math.isfinite(out["loss_mc_last"])
    # first step: same weights, same data, same seeds -> the Monte-Carlo loss agrees to fp32 noise
    assert abs(out["loss_mc_first"] - base["loss_mc_first"]) < 1e-3 * max(1.0, abs(base["loss_mc_first"]))
