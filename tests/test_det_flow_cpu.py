"""The call sequence of the detection head (EPro-PnP-Det/epropnp_det/models/dense_heads/deform_pnp_head.py:
test_post :514-549, loss :871-893) on the drop-in package, 4DoF, with the kernels running under the CPU SIMT emulator:
config-built layer, camera.set_param(img_shape=...), adaptive cost, fast-mode Monte-Carlo inference with the
random-sample initialiser, orientation-grid evaluate_pnp, training-mode Monte-Carlo forward + backward, and the
derivative-regularisation branch (with_pose_opt_plus) + backward.  Small sizes: RSLM launches one CTA per proposal."""
import math

import pytest
import torch

import simt_native
from epropnp.builder import build_camera, build_cost_fun, build_pnp
from epropnp.common import evaluate_pnp
from epropnp_b200.synth import make_problem
from oracle import pnp_oracle as orc


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


def _layer():
    """The layer exactly as EPro-PnP-Det/configs/epropnp_det_basic.py:98-111 builds it (fewer samples / proposals only)."""
    return build_pnp(dict(type="EProPnP4DoF", mc_samples=128, num_iter=4, normalize=True,
                          solver=dict(type="LMSolver", num_iter=10, normalize=True,
                                      init_solver=dict(type="RSLMSolver", num_points=16, num_proposals=16, num_iter=3))))


def test_inference_sequence(dev, monkeypatch):
    B, N = 3, 48
    pc = make_problem(B, N, seed=21, dof=4)
    x3d, x2d, w2d, gt = pc["x3d"], pc["x2d"], pc["w2d"], pc["pose_gt"]
    pnp, camera, cost_fun = _layer(), build_camera(dict(type="PerspectiveCamera", z_min=0.5)), \
        build_cost_fun(dict(type="AdaptiveHuberPnPCost", relative_delta=0.5))
    ori_shapes = torch.tensor([[480.0, 640.0]]).expand(B, 2)                      # (h, w) per object
    camera.set_param(pc["cam_mats"], img_shape=ori_shapes)
    assert camera.lb == -200.5 and camera.ub.shape == (B, 2)      # scalar lower bound, per-object upper bound (camera.py:55-59)
    cost_fun.set_param(x2d.detach(), w2d)
    torch.manual_seed(3)
    with torch.no_grad():
        pose_opt, _, _, samples, logw, _ = pnp.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, fast_mode=True)
        pose_only = pnp(x3d, x2d, w2d, camera, cost_fun, fast_mode=True)[0]
    assert pose_opt.shape == (B, 4) and samples.shape == (128, B, 4) and logw.shape == (128, B)
    assert torch.isfinite(logw).all() and pose_only.shape == (B, 4)
    weights = logw.softmax(dim=0)                                                 # head: score from sample spread
    dev_te = (samples[..., [0, 2]] - pose_opt[:, [0, 2]]).norm(dim=-1)
    assert torch.isfinite((dev_te * weights).sum(dim=0)).all()
    # the head's Monte-Carlo score (deform_pnp_head.py:533-536) through the package: torch composite == native epilogue
    from epropnp import monte_carlo_pose_loss as mcl
    score_ref = (((-dev_te.log2() + 2.5) / 4).clamp(min=0, max=1) * weights).sum(dim=0)
    assert torch.allclose(mcl.mc_score_te(samples, pose_opt, logw), score_ref, atol=1e-6)
    monkeypatch.setattr(mcl, "_use_native", lambda t: True)
    assert torch.allclose(mcl.mc_score_te(samples, pose_opt, logw), score_ref, atol=2e-6)
    assert torch.allclose(mcl.mc_sample_weights(logw), weights, atol=1e-6)
    # with 8 random-sample proposals most objects land on the ground truth
    good = ((pose_opt[:, :3] - gt[:, :3]).norm(dim=-1) < 0.2).float().mean()
    assert good >= 2 / 3
    # orientation grid (head :538-551): cost of (bins, B, 4) poses against the oracle
    bins = 16
    grid = pose_opt[None].expand(bins, -1, -1).clone()
    grid[..., 3] = torch.linspace(0, 2 * math.pi * (bins - 1) / bins, bins)[:, None]
    cost = evaluate_pnp(x3d, x2d, w2d, grid, camera, cost_fun, out_cost=True)[1]
    ref = orc.evaluate(x3d.double(), x2d.double(), w2d.double(), grid.double(),
                       orc.Camera(pc["cam_mats"].double(), 0.5, torch.full((B, 2), camera.lb, dtype=torch.float64), camera.ub.double()),
                       cost_fun.delta.double())["cost"]
    assert cost.shape == (bins, B) and torch.allclose(cost.double(), ref, rtol=2e-5, atol=1e-4)
    assert cost.neg().log_softmax(dim=0).transpose(1, 0).shape == (B, bins)


def test_training_sequence(dev):
    B, N = 4, 32
    pc = make_problem(B, N, seed=22, dof=4)
    x2d, gt = pc["x2d"], pc["pose_gt"]
    x3d = pc["x3d"].clone().requires_grad_(True)
    w2d = pc["w2d"].clone().requires_grad_(True)
    scale = torch.full((B, 1, 2), 1.5, requires_grad=True)
    pnp, camera, cost_fun = _layer(), build_camera(dict(type="PerspectiveCamera", z_min=0.5)), \
        build_cost_fun(dict(type="AdaptiveHuberPnPCost", relative_delta=0.5))
    camera.set_param(pc["cam_mats"], img_shape=torch.tensor([[480.0, 640.0]]).expand(B, 2))
    w2d_scaled = w2d * scale
    cost_fun.set_param(x2d.detach(), w2d_scaled)
    torch.manual_seed(4)
    _, _, _, _, logw, cost_tgt = pnp.monte_carlo_forward(x3d, x2d, w2d_scaled, camera, cost_fun,
                                                         pose_init=gt, force_init_solve=True)
    assert logw.shape == (128, B) and cost_tgt.shape == (B,) and logw.requires_grad and cost_tgt.requires_grad
    loss_pose = (cost_tgt + torch.logsumexp(logw, dim=0)).mean()                  # Monte-Carlo pose loss
    loss_pose.backward()
    for t in (x3d, w2d, scale):
        assert t.grad is not None and torch.isfinite(t.grad).all() and t.grad.abs().sum() > 0
    # derivative regularisation (head :886-893): detached delta, pose_opt_plus differentiable
    x3d.grad = w2d.grad = None
    cost_fun.delta = cost_fun.delta.detach()
    pose_opt, _, _, pose_opt_plus = pnp(x3d, x2d, w2d * scale.detach(), camera, cost_fun, with_pose_opt_plus=True)
    assert not pose_opt.requires_grad and pose_opt_plus.requires_grad
    loss_reg = (pose_opt_plus[:, :3] - gt[:, :3]).norm(dim=-1).mean()
    loss_reg.backward()
    assert torch.isfinite(x3d.grad).all() and x3d.grad.abs().sum() > 0 and torch.isfinite(w2d.grad).all()
