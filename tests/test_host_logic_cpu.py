"""Host logic of the drop-in package on CPU: the tensor-level native entry points are swapped for an oracle-backed
fake (tests/fake_native.py, test-only), everything above them -- argument plumbing of camera / cost objects, shapes and
(M, B) views, normalise / denormalise, the RSLM initialiser's bookkeeping, the autograd bridge -- is the product's own
code, checked in float64 against golden vectors from the unmodified reference (values AND gradients)."""
import numpy as np
import pytest
import torch

import fake_native
from conftest import err_vs, golden_bounds, golden_names, load_golden
from epropnp.camera import PerspectiveCamera
from epropnp.common import evaluate_pnp, pnp_denormalize, pnp_normalize
from epropnp.cost_fun import AdaptiveHuberPnPCost, HuberPnPCost
from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
from epropnp.levenberg_marquardt import LMSolver, RSLMSolver
from epropnp_b200.synth import make_problem

D = torch.float64


@pytest.fixture(autouse=True)
def _fake(monkeypatch):
    fake_native.install(monkeypatch)


def _setup(g, grad=False):
    t = lambda k: torch.from_numpy(g[k]).to(D)
    x3d, x2d, w2d = t("x3d"), t("x2d"), t("w2d")
    if grad:
        x3d, x2d, w2d = (v.requires_grad_(True) for v in (x3d, x2d, w2d))
    lb, ub = golden_bounds(g, D)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]), lb=lb, ub=ub)
    if float(g["fixed_delta"]) >= 0:
        cost_fun = HuberPnPCost(delta=float(g["fixed_delta"]))
    else:
        cost_fun = AdaptiveHuberPnPCost(relative_delta=float(g["relative_delta"]))
        cost_fun.set_param(x2d.detach(), w2d)
    return x3d, x2d, w2d, camera, cost_fun, t("pose_init")


def _noise(g, key="noise_rot"):
    B = int(g["B"])
    n3 = torch.from_numpy(np.transpose(g["noise_normal"], (2, 0, 1, 3)).reshape(B, -1, 3).copy()).to(D)
    c2 = torch.from_numpy(np.transpose(g["noise_chi2"], (2, 0, 1)).reshape(B, -1).copy()).to(D)
    r = g[key]
    r = np.transpose(r, (2, 0, 1, 3)).reshape(B, -1, 4) if r.ndim == 4 else np.transpose(r, (2, 0, 1)).reshape(B, -1)
    return n3, c2, torch.from_numpy(r.copy()).to(D)


@pytest.mark.parametrize("name", golden_names("lm") + golden_names("gn"))
def test_lmsolver_plumbing(name):
    g = load_golden(name)
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g)
    solver = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]))
    pose, cov, cost, plus = solver(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_cov=True,
                                   with_cost=True, with_pose_opt_plus=True, fast_mode=bool(g["fast_mode"]))
    assert err_vs(pose, g["ref64_lm_pose"]) < 1e-8 and err_vs(cov, g["ref64_lm_cov"]) < 1e-6
    assert err_vs(cost, g["ref64_lm_cost"]) < 1e-8 and err_vs(plus, g["ref64_lm_pose_plus"]) < 1e-8
    if int(g["normalize"]):
        out = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]), normalize=True)(
            x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_cost=True)
        assert err_vs(out[0], g["ref64_lmnorm_pose"]) < 1e-8 and out[1] is None
        via_override = solver(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, normalize_override=True)
        assert err_vs(via_override[0], g["ref64_lmnorm_pose"]) < 1e-8
    res, c, jac = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun, out_jacobian=True, out_residual=True,
                               out_cost=True, clip_jac=not bool(g["fast_mode"]))
    assert err_vs(jac, g["ref64_eval_jac"]) < 1e-9 and err_vs(res, g["ref64_eval_residual"]) < 1e-9
    cm = evaluate_pnp(x3d, x2d, w2d, torch.from_numpy(g["eval_poses"]).to(D), camera, cost_fun, out_cost=True)[1]
    assert err_vs(cm, g["ref64_eval_cost_multi"]) < 1e-9


@pytest.mark.parametrize("name", golden_names("mc"))
def test_monte_carlo_forward_plumbing(name):
    g = load_golden(name)
    dof = int(g["dof"])
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    layer = (EProPnP6DoF if dof == 6 else EProPnP4DoF)(mc_samples=M, num_iter=I,
                                                        solver=LMSolver(dof=dof, num_iter=int(g["lm_iter"])))
    noise = _noise(g, "noise_rot" if dof == 6 else "yaw_samples64")
    pose_opt, cost, plus, samples, logw, cost_init = layer.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False, with_cost=True, amis_noise=noise)
    assert samples.shape == g["ref64_mc_samples"].shape and logw.shape == g["ref64_mc_logw"].shape and plus is None
    assert err_vs(pose_opt, g["ref64_mc_pose"]) < 1e-8 and err_vs(cost, g["ref64_mc_cost"]) < 1e-8
    assert err_vs(samples, g["ref64_mc_samples"]) < 1e-7 and err_vs(logw, g["ref64_mc_logw"]) < 1e-7
    assert err_vs(cost_init, g["ref64_mc_cost_init"]) < 1e-9


def test_monte_carlo_forward_with_normalize():
    """normalize=True (all EPro-PnP-Det configs): samples / poses come back in the original frame."""
    from oracle import pnp_oracle as orc
    g = load_golden("mc6_basic")
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    noise = _noise(g)
    layer = EProPnP6DoF(mc_samples=M, num_iter=I, normalize=True, solver=LMSolver(dof=6, num_iter=10))
    pose_opt, _, plus, samples, logw, cost_init = layer.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False, with_pose_opt_plus=True,
        amis_noise=noise)
    off, x3n, p0n = orc.normalize_points(x3d, pose_init)
    S = M // I
    nz = (noise[0].reshape(-1, I, S, 3).permute(1, 2, 0, 3), noise[1].reshape(-1, I, S).permute(1, 2, 0),
          noise[2].reshape(-1, I, S, 4).permute(1, 2, 0, 3))
    ref = orc.monte_carlo_forward_6dof(x3n, x2d, w2d, orc.Camera(camera.cam_mats, camera.z_min), cost_fun.delta, p0n, nz, M, I)
    assert err_vs(pose_opt, orc.denormalize_pose(off, ref["pose_opt"])) < 1e-8
    assert err_vs(samples, orc.denormalize_pose(off, ref["samples"])) < 1e-8
    assert err_vs(logw, ref["logw"]) < 1e-7 and err_vs(cost_init, ref["cost_init"]) < 1e-9
    assert plus.shape == pose_opt.shape and torch.isfinite(plus).all()
    # helpers on stacked poses
    o2, xn2, _ = pnp_normalize(x3d)
    assert torch.allclose(pnp_denormalize(o2, ref["samples"]), orc.denormalize_pose(off, ref["samples"]))


def test_rslm_and_force_init_solve_bookkeeping():
    B, N = 5, 48
    pc = make_problem(B, N, seed=4)
    x3d, x2d, w2d = (pc[k].to(D) for k in ("x3d", "x2d", "w2d"))
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"].to(D))
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(x2d, w2d)
    torch.manual_seed(0)
    rs = RSLMSolver(dof=6, num_points=8, num_proposals=48, num_iter=5)
    pose, none, min_cost = rs.solve(x3d, x2d, w2d, camera, cost_fun, with_cost=True)
    assert pose.shape == (B, 7) and none is None and min_cost.shape == (B,)
    full = evaluate_pnp(x3d, x2d, w2d, pose, camera, cost_fun, out_cost=True)[1]
    assert torch.allclose(full, min_cost)                      # the returned cost is the full-set cost of the winner
    solver = LMSolver(dof=6, num_iter=10, init_solver=rs)
    gt = pc["pose_gt"].to(D)
    p1 = solver(x3d, x2d, w2d, camera, cost_fun, with_cost=True)                  # pose_init=None -> RSLM start
    assert ((p1[0][:, :3] - gt[:, :3]).norm(dim=-1) < 0.15).float().mean() >= 0.8
    # force_init_solve keeps the given pose where it is better than the random-sample solution
    good = gt.clone()
    p2 = solver(x3d, x2d, w2d, camera, cost_fun, pose_init=good, force_init_solve=True, with_cost=True)
    c_good = evaluate_pnp(x3d, x2d, w2d, good, camera, cost_fun, out_cost=True)[1]
    assert (p2[2] <= c_good + 1e-9).all()
    # 4DoF center-based init uses the y-extent ratio (levenberg_marquardt.py:289-292)
    t4 = RSLMSolver(dof=4).center_based_init(x2d, x3d, camera)
    t6 = rs.center_based_init(x2d, x3d, camera)
    assert t4.shape == t6.shape == (B, 3) and (t6[:, 2] > 0).all()
    with pytest.raises(AssertionError):
        LMSolver(dof=6)(x3d, x2d, w2d, camera, cost_fun)       # no pose_init and no init_solver


@pytest.mark.parametrize("name", ["mc6_basic", "mc6_bounds", "mc4_basic"])
def test_autograd_bridge_against_reference_gradients(name):
    """Signs, transposes, the delta -> w2d chain and the GN-step composite of epropnp/autograd.py, in float64 with the
    reference's own noise: gradients must equal the unmodified reference's autograd."""
    g = load_golden(name)
    dof = int(g["dof"])
    x3d, x2d, w2d, camera, cost_fun, pose_init = _setup(g, grad=True)
    assert cost_fun.delta.requires_grad
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    layer = (EProPnP6DoF if dof == 6 else EProPnP4DoF)(mc_samples=M, num_iter=I,
                                                        solver=LMSolver(dof=dof, num_iter=int(g["lm_iter"])))
    noise = _noise(g, "noise_rot" if dof == 6 else "yaw_samples64")
    pose_opt, cost, plus, samples, logw, cost_init = layer.monte_carlo_forward(
        x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, force_init_solve=False, with_pose_opt_plus=True,
        amis_noise=noise)
    assert logw.requires_grad and cost_init.requires_grad and plus.requires_grad and not samples.requires_grad
    assert err_vs(logw.detach(), g["ref64_grad_logw"]) < 1e-7 and err_vs(plus.detach(), g["ref64_grad_pose_plus"]) < 1e-8
    t = lambda k: torch.from_numpy(g[k]).to(D)
    g1 = torch.autograd.grad((t("grad_c1") * logw).sum() + (t("grad_c2") * cost_init).sum(), [x3d, x2d, w2d], retain_graph=True)
    g2 = torch.autograd.grad((t("grad_c3") * plus).sum(), [x3d, x2d, w2d])
    for nm, a, b in zip(("x3d", "x2d", "w2d"), g1, g2):
        assert err_vs(a, g[f"ref64_gradL1_{nm}"]) < 1e-6, ("L1", nm)
        assert err_vs(b, g[f"ref64_gradL2_{nm}"]) < 1e-6, ("L2", nm)
    # LMSolver.forward alone: differentiable pose_opt_plus
    solver = LMSolver(dof=dof, num_iter=int(g["lm_iter"]))
    out = solver(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_opt_plus=True)
    assert out[3].requires_grad and not out[0].requires_grad
    # evaluate_pnp cost with gradients
    c = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun, out_cost=True)[1]
    assert c.requires_grad and err_vs(c.detach(), g["ref64_grad_cost_init"]) < 1e-9


# ------------------------------------------------------------------------------------------------ config builders
def test_builders_construct_nested_configs():
    """Detection-style configs (EPro-PnP-Det/configs/*: dicts with `type`, nested solver / init_solver)."""
    from epropnp.builder import CAMERA, COSTFUN, PNP, build_camera, build_cost_fun, build_pnp
    from epropnp.camera import PerspectiveCamera
    from epropnp.cost_fun import AdaptiveHuberPnPCost
    from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
    from epropnp.levenberg_marquardt import LMSolver, RSLMSolver
    assert {"LMSolver", "RSLMSolver", "EProPnP4DoF", "EProPnP6DoF"} <= set(PNP.module_dict)
    assert "PerspectiveCamera" in CAMERA and {"HuberPnPCost", "AdaptiveHuberPnPCost"} <= set(COSTFUN.module_dict)
    pnp = build_pnp(dict(type="EProPnP4DoF", mc_samples=512, num_iter=4,
                         solver=dict(type="LMSolver", dof=4, num_iter=5,
                                     init_solver=dict(type="RSLMSolver", dof=4, num_points=16, num_proposals=64, num_iter=3))))
    assert isinstance(pnp, EProPnP4DoF) and isinstance(pnp.solver, LMSolver) and pnp.solver.num_iter == 5
    assert isinstance(pnp.solver.init_solver, RSLMSolver) and pnp.solver.init_solver.num_proposals == 64
    assert pnp.iter_samples == 128
    cam = build_camera(dict(type="PerspectiveCamera"), z_min=0.5)
    assert isinstance(cam, PerspectiveCamera) and cam.z_min == 0.5
    cost = build_cost_fun(dict(type="AdaptiveHuberPnPCost", relative_delta=0.5))
    assert isinstance(cost, AdaptiveHuberPnPCost)
    # instances pass through (canonical / 6DoF style), unknown names and missing `type` are errors
    solver = LMSolver(dof=6)
    assert EProPnP6DoF(solver=solver).solver is solver
    assert build_pnp(dict(type=EProPnP6DoF, mc_samples=64, num_iter=2, solver=solver)).mc_samples == 64
    with pytest.raises(KeyError):
        build_pnp(dict(type="NoSuchSolver"))
    with pytest.raises(KeyError):
        build_pnp(dict(dof=4))
    with pytest.raises(KeyError):
        PNP.register_module(module=LMSolver)          # duplicate registration


def test_empty_batch_stays_in_the_autograd_graph():
    """B = 0 (a detection rank without objects): the layer's differentiable outputs must carry a grad_fn like the
    reference's (epropnp.py:184-187), so that a loss made only of them can be back-propagated (zero gradients)."""
    import torch
    from epropnp.camera import PerspectiveCamera
    from epropnp.cost_fun import AdaptiveHuberPnPCost
    from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
    from epropnp.levenberg_marquardt import LMSolver
    from epropnp.monte_carlo_pose_loss import MonteCarloPoseLoss
    for cls, dof, pd in ((EProPnP6DoF, 6, 7), (EProPnP4DoF, 4, 4)):
        x3d = torch.zeros(0, 16, 3, requires_grad=True)
        x2d = torch.zeros(0, 16, 2, requires_grad=True)
        w2d = torch.zeros(0, 16, 2, requires_grad=True)
        camera = PerspectiveCamera(cam_mats=torch.zeros(0, 3, 3))
        cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
        cost_fun.set_param(x2d.detach(), w2d)
        layer = cls(mc_samples=8, num_iter=2, solver=LMSolver(dof=dof, num_iter=2))
        pose_opt, cost, plus, samples, logw, cost_init = layer.monte_carlo_forward(
            x3d, x2d, w2d, camera, cost_fun, pose_init=torch.zeros(0, pd), force_init_solve=False, with_pose_opt_plus=True)
        assert pose_opt.shape == (0, pd) and samples.shape == (8, 0, pd) and logw.shape == (8, 0)
        assert logw.grad_fn is not None and cost_init.grad_fn is not None and plus.grad_fn is not None
        loss = MonteCarloPoseLoss()(logw, cost_init, 1.0) + plus.sum()
        loss.backward()
        for t in (x3d, x2d, w2d):
            assert t.grad is not None and t.grad.shape == t.shape


def test_evaluate_pnp_differentiable_outputs_go_through_the_composite(monkeypatch):
    """evaluate_pnp with grad-requiring inputs and a residual / Jacobian request (or a grad-requiring pose): the torch
    composite of camera.project + cost_fun.compute, as LMSolver.gn_step needs under autograd (reference common.py:67-100)
    -- values equal the oracle's, and gradients flow to x3d and to the pose."""
    import torch
    pc = make_problem(3, 24, seed=9)
    x3d = pc["x3d"].to(D).requires_grad_(True)
    x2d, w2d = pc["x2d"].to(D), pc["w2d"].to(D)
    pose = pc["pose_init"].to(D).requires_grad_(True)
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"].to(D))
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(x2d, w2d)
    res, cost, jac = evaluate_pnp(x3d, x2d, w2d, pose, camera, cost_fun, out_jacobian=True, out_residual=True, out_cost=True)
    from oracle import pnp_oracle as orc
    e = orc.evaluate(x3d.detach(), x2d, w2d, pose.detach(), orc.Camera(pc["cam_mats"].to(D), 0.1), cost_fun.delta.detach(), want_jac=True)
    assert torch.allclose(res, e["residual"], atol=1e-9) and torch.allclose(jac, e["jac"], atol=1e-9)
    assert torch.allclose(cost, e["cost"], atol=1e-9)
    (cost.sum() + res.square().sum()).backward()
    assert x3d.grad is not None and pose.grad is not None and x3d.grad.abs().sum() > 0 and pose.grad.abs().sum() > 0
