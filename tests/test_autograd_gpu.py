"""Training path on the GPU: forward = fused kernel, backward = native Monte-Carlo cost gradient
(epnp_cost_backward_f32), against gradients recorded from the unmodified reference's autograd (float64)."""
import numpy as np
import pytest
import torch

from conftest import err_vs, golden_bounds, load_golden
from epropnp.camera import PerspectiveCamera
from epropnp.common import evaluate_pnp
from epropnp.cost_fun import AdaptiveHuberPnPCost
from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
from epropnp.levenberg_marquardt import LMSolver
from epropnp_b200 import native
from epropnp_b200.synth import make_problem

pytestmark = pytest.mark.gpu


def _noise(g, dev, dof):
    B = int(g["B"])
    n3 = torch.from_numpy(np.transpose(g["noise_normal"], (2, 0, 1, 3)).reshape(B, -1, 3).copy()).to(dev)
    c2 = torch.from_numpy(np.transpose(g["noise_chi2"], (2, 0, 1)).reshape(B, -1).copy()).to(dev)
    if dof == 6:
        r = torch.from_numpy(np.transpose(g["noise_rot"], (2, 0, 1, 3)).reshape(B, -1, 4).copy()).to(dev)
    else:
        r = torch.from_numpy(np.transpose(g["yaw_samples64"], (2, 0, 1)).reshape(B, -1).copy()).float().to(dev)
    return n3, c2, r


@pytest.mark.parametrize("name", ["mc6_basic", "mc6_bounds", "mc4_basic"])
def test_monte_carlo_backward_matches_reference(cuda_device, name):
    g = load_golden(name)
    dev, dof = cuda_device, int(g["dof"])
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    x3d, x2d, w2d = (t(k).float().requires_grad_(True) for k in ("x3d", "x2d", "w2d"))
    lb, ub = golden_bounds(g)
    if torch.is_tensor(lb):
        lb, ub = lb.to(dev), ub.to(dev)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]), lb=lb, ub=ub)
    cost_fun = AdaptiveHuberPnPCost(relative_delta=float(g["relative_delta"]))
    cost_fun.set_param(x2d.detach(), w2d)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    cls = EProPnP6DoF if dof == 6 else EProPnP4DoF
    layer = cls(mc_samples=M, num_iter=I, solver=LMSolver(dof=dof, num_iter=int(g["lm_iter"])))
    pose_opt, cost, plus, samples, logw, cost_init = layer.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=t("pose_init"), force_init_solve=False, with_pose_opt_plus=True,
        amis_noise=_noise(g, dev, dof))
    assert logw.requires_grad and cost_init.requires_grad and plus.requires_grad and not samples.requires_grad
    assert logw.shape == (M, int(g["B"])) and err_vs(cost_init.detach().cpu(), g["ref64_grad_cost_init"]) < 2e-5
    c1, c2, c3 = t("grad_c1").float(), t("grad_c2").float(), t("grad_c3").float()
    g1 = torch.autograd.grad((c1 * logw).sum() + (c2 * cost_init).sum(), [x3d, x2d, w2d], retain_graph=True)
    g2 = torch.autograd.grad((c3 * plus).sum(), [x3d, x2d, w2d])
    # fp32 forward (samples differ from the fp64 reference's by its fp32 floor) -> 2e-3 of the gradient scale
    for nm, a, b in zip(("x3d", "x2d", "w2d"), g1, g2):
        assert err_vs(a.cpu(), g[f"ref64_gradL1_{nm}"]) < 2e-3, ("L1", nm)
        assert err_vs(b.cpu(), g[f"ref64_gradL2_{nm}"]) < 5e-3, ("L2", nm)


def test_cost_backward_kernel_against_oracle_autograd(cuda_device):
    """Native backward at the north-star shape (N = 512, 513 poses per object) vs fp64 autograd of the oracle."""
    from oracle import pnp_oracle as orc
    B, N, P = 6, 512, 513
    pc = make_problem(B, N, seed=31, outlier_frac=0.1)
    dev = cuda_device
    gen = torch.Generator().manual_seed(2)
    poses = pc["pose_gt"][:, None, :].repeat(1, P, 1)
    poses[..., :3] += 0.05 * torch.randn(B, P, 3, generator=gen)
    q = poses[..., 3:] + 0.03 * torch.randn(B, P, 4, generator=gen)
    poses[..., 3:] = q / q.norm(dim=-1, keepdim=True)
    up = torch.randn(B, P, generator=gen)
    delta = orc.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    prob = native.Problem(pc["x3d"].to(dev), pc["x2d"].to(dev), pc["w2d"].to(dev), pc["cam_mats"].to(dev), None, None,
                          delta.to(dev))
    gx3d, gx2d, gw2d, gdel = native.cost_backward(prob, 6, 0.1, poses[:, :P - 1].to(dev), up[:, :P - 1].to(dev),
                                                  poses[:, P - 1:].to(dev), up[:, P - 1:].to(dev))
    d = torch.float64
    t3, t2, tw = (pc[k].to(d).requires_grad_(True) for k in ("x3d", "x2d", "w2d"))
    td = delta.to(d).requires_grad_(True)
    cost = orc.evaluate(t3, t2, tw, poses.transpose(0, 1).to(d), orc.Camera(pc["cam_mats"].to(d), 0.1), td)["cost"]
    (cost * up.T.to(d)).sum().backward()
    assert err_vs(gx3d.cpu(), t3.grad) < 2e-4 and err_vs(gx2d.cpu(), t2.grad) < 2e-4
    assert err_vs(gw2d.cpu(), tw.grad) < 2e-4 and err_vs(gdel.cpu(), td.grad) < 2e-4


def test_evaluate_pnp_cost_is_differentiable(cuda_device):
    g = load_golden("lm6_basic")        # smooth case (no clamps hit): finite differences are meaningful
    dev = cuda_device
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    x3d = t("x3d").requires_grad_(True)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]))
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(t("x2d"), t("w2d"))
    poses = t("eval_poses")
    cost = evaluate_pnp(x3d, t("x2d"), t("w2d"), poses, camera, cost_fun, out_cost=True)[1]
    assert cost.shape == poses.shape[:2] and cost.requires_grad
    cost.sum().backward()
    # finite difference along a random direction
    v = torch.randn_like(x3d)
    eps = 1e-3
    with torch.no_grad():
        cp = evaluate_pnp(x3d + eps * v, t("x2d"), t("w2d"), poses, camera, cost_fun, out_cost=True)[1].double().sum()
        cm = evaluate_pnp(x3d - eps * v, t("x2d"), t("w2d"), poses, camera, cost_fun, out_cost=True)[1].double().sum()
    fd = (cp - cm) / (2 * eps)
    an = (x3d.grad.double() * v.double()).sum()
    assert abs(fd - an) / abs(an) < 3e-2
    # gradients with respect to the POSE: the torch composite (reference common.py:67-100 is differentiable there too)
    pg = poses.clone().requires_grad_(True)
    c2 = evaluate_pnp(x3d.detach(), t("x2d"), t("w2d"), pg, camera, cost_fun, out_cost=True)[1]
    assert torch.allclose(c2.detach(), cost.detach(), rtol=2e-4, atol=1e-3)
    c2.sum().backward()
    assert pg.grad is not None and torch.isfinite(pg.grad).all() and pg.grad.abs().sum() > 0
