"""The reference-facing Python surface (epropnp.*) on CUDA tensors: same calls a user of the reference
makes, results against the reference's golden vectors.  Reads like the reference's own usage
(demo/fit_identity.ipynb cell 7/10, EPro-PnP-6DoF/lib/test.py:200-221)."""
import numpy as np
import pytest
import torch

from conftest import assert_lm_parity, err_vs, golden_bounds, golden_names, load_golden
from epropnp.camera import PerspectiveCamera
from epropnp.common import evaluate_pnp
from epropnp.cost_fun import AdaptiveHuberPnPCost, HuberPnPCost
from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
from epropnp.levenberg_marquardt import LMSolver, RSLMSolver
from epropnp_b200.synth import make_problem

pytestmark = pytest.mark.gpu


def _setup(g, dev):
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    lb, ub = golden_bounds(g)
    if torch.is_tensor(lb):
        lb, ub = lb.to(dev), ub.to(dev)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]), lb=lb, ub=ub)
    if float(g["fixed_delta"]) >= 0:
        cost_fun = HuberPnPCost(delta=float(g["fixed_delta"]))
    else:
        cost_fun = AdaptiveHuberPnPCost(relative_delta=float(g["relative_delta"]))
        cost_fun.set_param(t("x2d"), t("w2d"))
    return t("x3d"), t("x2d"), t("w2d"), camera, cost_fun, t("pose_init")


@pytest.mark.parametrize("name", golden_names("lm") + golden_names("gn"))
def test_lmsolver_forward(cuda_device, name):
    g = load_golden(name)
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g, cuda_device)
    solver = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]))
    pose, cov, cost, plus = solver(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_cov=True,
                                   with_cost=True, with_pose_opt_plus=True, fast_mode=bool(g["fast_mode"]))
    tol = max(1e-4, 3 * err_vs(g["ref32_lm_pose"], g["ref64_lm_pose"]))
    assert_lm_parity(pose.cpu().numpy(), cost.cpu().numpy(), g["ref32_lm_pose"], g["ref64_lm_cost"], tol, what=name)
    assert_lm_parity(plus.cpu().numpy(), cost.cpu().numpy(), g["ref32_lm_pose_plus"], g["ref64_lm_cost"], tol, what=name)
    assert cov.shape == g["ref32_lm_cov"].shape and cost.shape == g["ref32_lm_cost"].shape
    pose2, cov2, cost2 = solver.solve(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init,
                                      fast_mode=bool(g["fast_mode"]))
    assert cov2 is None and cost2 is None and torch.equal(pose2, pose)
    if int(g["normalize"]):
        sn = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]), normalize=True)
        pn = sn(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_cost=True)
        assert_lm_parity(pn[0].cpu().numpy(), pn[2].cpu().numpy(), g["ref32_lmnorm_pose"], g["ref64_lm_cost"], tol, what=name)
        with pytest.raises(NotImplementedError):
            sn(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_cov=True)


def test_evaluate_pnp_semantics(cuda_device):
    g = load_golden("lm6_bounds")
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g, cuda_device)
    B, N = x2d.shape[:2]
    res, cost, jac = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun, out_jacobian=True, out_residual=True,
                                  out_cost=True)
    assert res.shape == (B, 2 * N) and jac.shape == (B, 2 * N, 6) and cost.shape == (B,)
    assert err_vs(jac.cpu().numpy(), g["ref32_eval_jac"]) < 2e-5
    # preallocated outputs are written in place (levenberg_marquardt.py:132-134 usage)
    jac_buf = torch.empty(B, 2 * N, 6, device=cuda_device)
    cost_buf = torch.empty(B, device=cuda_device)
    r2, c2, j2 = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun, out_jacobian=jac_buf, out_cost=cost_buf)
    assert r2 is None and j2.data_ptr() == jac_buf.data_ptr() and torch.equal(jac_buf, jac) and torch.equal(cost_buf, cost)
    # stacked hypotheses (S, B, D) broadcast against (B, N, .)   (deform_pnp_head.py:548 usage)
    poses = torch.from_numpy(g["eval_poses"]).to(cuda_device)
    c = evaluate_pnp(x3d, x2d, w2d, poses, camera, cost_fun, out_cost=True)[1]
    assert c.shape == poses.shape[:2] and err_vs(c.cpu().numpy(), g["ref32_eval_cost_multi"]) < 2e-5
    # clip_jac=False keeps gradients of clamped points
    j_noclip = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun, out_jacobian=True, clip_jac=False)[2]
    assert (j_noclip != 0).sum() > (jac != 0).sum()
    # grad-requiring inputs + a Jacobian request: the torch composite (what LMSolver.gn_step needs under autograd)
    xg = x3d.clone().requires_grad_(True)
    rg, _, jg = evaluate_pnp(xg, x2d, w2d, pose_init, camera, cost_fun, out_jacobian=True, out_residual=True)
    assert jg.grad_fn is not None and err_vs(jg.detach().cpu().numpy(), jac.cpu().numpy()) < 1e-5
    rg.square().sum().backward()
    assert xg.grad is not None and xg.grad.abs().sum() > 0


@pytest.mark.parametrize("name", golden_names("mc6"))
def test_monte_carlo_forward(cuda_device, name):
    g = load_golden(name)
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g, cuda_device)
    B, M, I = int(g["B"]), int(g["mc_samples_total"]), int(g["mc_iters"])
    layer = EProPnP6DoF(mc_samples=M, num_iter=I, solver=LMSolver(dof=6, num_iter=int(g["lm_iter"])))
    noise = (torch.from_numpy(np.transpose(g["noise_normal"], (2, 0, 1, 3)).reshape(B, -1, 3).copy()).to(cuda_device),
             torch.from_numpy(np.transpose(g["noise_chi2"], (2, 0, 1)).reshape(B, -1).copy()).to(cuda_device),
             torch.from_numpy(np.transpose(g["noise_rot"], (2, 0, 1, 3)).reshape(B, -1, 4).copy()).to(cuda_device))
    pose_opt, cost, plus, samples, logw, cost_init = layer.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False, with_cost=True, amis_noise=noise)
    assert samples.shape == (M, B, 7) and logw.shape == (M, B) and plus is None
    floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
    assert err_vs(logw.cpu().numpy(), g["ref32_mc_logw"]) < max(1e-4, 5 * floor_w)
    assert err_vs(pose_opt.cpu().numpy(), g["ref32_mc_pose"]) < max(1e-4, 3 * err_vs(g["ref32_mc_pose"], g["ref64_mc_pose"]))
    assert err_vs(cost_init.cpu().numpy(), g["ref32_mc_cost_init"]) < 2e-5
    # reference call sites normalise over dim 0 (EPro-PnP-6DoF/lib/models/monte_carlo_pose_loss.py:29)
    assert torch.allclose(torch.softmax(logw, dim=0).sum(0), torch.ones(B, device=cuda_device), atol=1e-5)
    # seeds: torch.manual_seed governs the production RNG
    torch.manual_seed(5)
    a = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False)
    torch.manual_seed(5)
    b = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False)
    assert torch.equal(a[4], b[4]) and a[1] is None


def test_layer_forward_and_4dof(cuda_device):
    g = load_golden("lm4_basic")
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g, cuda_device)
    layer = EProPnP4DoF(mc_samples=512, num_iter=4, solver=LMSolver(dof=4, num_iter=10))
    pose, cov, cost, plus = layer(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_cov=True)
    tol = max(1e-4, 3 * err_vs(g["ref32_lm_pose"], g["ref64_lm_pose"]))
    assert err_vs(pose.cpu().numpy(), g["ref32_lm_pose"]) < tol and cov.shape == (int(g["B"]), 4, 4)
    r = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False,
                                  with_cost=True, amis_seed=11)
    assert r[3].shape == (512, int(g["B"]), 4) and r[4].shape == (512, int(g["B"])) and torch.isfinite(r[4]).all()
    assert torch.equal(r[0], pose)


def test_rslm_init_and_force_init_solve(cuda_device):
    """pose_init=None -> random-sample LM initialiser (demo notebook configuration, cell 7)."""
    B, N = 64, 64
    pc = make_problem(B, N, seed=9)
    dev = cuda_device
    x3d, x2d, w2d = pc["x3d"].to(dev), pc["x2d"].to(dev), pc["w2d"].to(dev)
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"].to(dev))
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(x2d, w2d)
    torch.manual_seed(0)
    solver = LMSolver(dof=6, num_iter=10, init_solver=RSLMSolver(dof=6, num_points=8, num_proposals=128, num_iter=5))
    pose, _, cost, _ = solver(x3d, x2d, w2d, camera, cost_fun, with_cost=True)
    gt = pc["pose_gt"].to(dev)
    dt = (pose[:, :3] - gt[:, :3]).norm(dim=-1)
    dq = 1 - (pose[:, 3:] * gt[:, 3:]).sum(-1).abs()
    assert ((dt < 0.1) & (dq < 1e-3)).float().mean() > 0.9          # recovers the pose without any prior
    layer = EProPnP6DoF(mc_samples=512, num_iter=4, solver=solver)
    bad_init = gt.clone()
    bad_init[:, :3] += 0.5
    r = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=bad_init, force_init_solve=True,
                                  with_cost=True)
    c_bad = evaluate_pnp(x3d, x2d, w2d, bad_init, camera, cost_fun, out_cost=True)[1]
    assert torch.allclose(r[5], c_bad) and (r[1] <= c_bad).all() and torch.isfinite(r[4]).all()
