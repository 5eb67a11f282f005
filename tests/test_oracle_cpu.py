"""Pins the oracle (oracle/pnp_oracle.py) against golden vectors produced by the UNMODIFIED
reference (oracle/make_golden.py).  float64: algorithmic identity (<=1e-8).  float32: rounding."""
import math

import numpy as np
import pytest
import torch

from conftest import err_vs, golden_bounds, golden_names, load_golden
from oracle import pnp_oracle as orc


def _setup(g, dtype):
    t = lambda k: torch.from_numpy(g[k]).to(dtype)
    lb, ub = golden_bounds(g, dtype)
    cam = orc.Camera(t("cam_mats"), float(g["z_min"]), lb, ub)
    delta = t("delta") if float(g["fixed_delta"]) < 0 else float(g["fixed_delta"])
    if float(g["fixed_delta"]) < 0:
        delta = orc.adaptive_delta(t("x2d"), t("w2d"), float(g["relative_delta"]))
    return t("x3d"), t("x2d"), t("w2d"), cam, delta, t("pose_init")


@pytest.mark.parametrize("name", golden_names())
@pytest.mark.parametrize("prec", ["ref64", "ref32"])
def test_evaluate_and_solve(name, prec):
    g = load_golden(name)
    dtype = torch.float64 if prec == "ref64" else torch.float32
    tol = 1e-8 if prec == "ref64" else 2e-4
    x3d, x2d, w2d, cam, delta, pose_init = _setup(g, dtype)
    if float(g["fixed_delta"]) < 0:
        assert err_vs(delta.numpy(), g["delta"]) < 1e-5
    fast = bool(g["fast_mode"])
    e = orc.evaluate(x3d, x2d, w2d, pose_init, cam, delta, want_jac=True, clip_jac=not fast)
    assert err_vs(e["cost"], g[prec + "_eval_cost"]) < tol
    assert err_vs(e["residual"], g[prec + "_eval_residual"]) < tol
    assert err_vs(e["jac"], g[prec + "_eval_jac"]) < tol
    poses = torch.from_numpy(g["eval_poses"]).to(dtype)
    c = orc.evaluate(x3d, x2d, w2d, poses, cam, delta)["cost"]
    assert err_vs(c, g[prec + "_eval_cost_multi"]) < tol

    prm = orc.LMParams(num_iter=int(g["lm_iter"]))
    pose, cov, cost = orc.lm_solve(x3d, x2d, w2d, cam, delta, pose_init, prm, fast_mode=fast)
    # float32: the reference itself moves by |ref32-ref64| when its rounding changes
    floor = err_vs(g["ref32_lm_pose"], g["ref64_lm_pose"])
    ptol = tol if prec == "ref64" else max(tol, 3 * floor)
    assert err_vs(pose, g[prec + "_lm_pose"]) < ptol
    assert err_vs(cost, g[prec + "_lm_cost"]) < max(ptol, 1e-5)
    assert err_vs(cov, g[prec + "_lm_cov"]) < (1e-6 if prec == "ref64" else 5e-3)
    plus = orc.pose_add(pose, orc.gn_step(x3d, x2d, w2d, cam, delta, pose))
    assert err_vs(plus, g[prec + "_lm_pose_plus"]) < ptol

    if int(g["normalize"]):
        off, x3n, p0n = orc.normalize_points(x3d, pose_init)
        pn, _, cn = orc.lm_solve(x3n, x2d, w2d, cam, delta, p0n, prm, fast_mode=fast)
        assert err_vs(orc.denormalize_pose(off, pn), g[prec + "_lmnorm_pose"]) < ptol


@pytest.mark.parametrize("name", golden_names("mc6"))
@pytest.mark.parametrize("prec", ["ref64", "ref32"])
def test_amis_6dof(name, prec):
    g = load_golden(name)
    dtype = torch.float64 if prec == "ref64" else torch.float32
    x3d, x2d, w2d, cam, delta, pose_init = _setup(g, dtype)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    noise = tuple(torch.from_numpy(g[k]).to(dtype) for k in ("noise_normal", "noise_chi2", "noise_rot"))
    r = orc.monte_carlo_forward_6dof(x3d, x2d, w2d, cam, delta, pose_init, noise, M, I,
                                     orc.LMParams(num_iter=int(g["lm_iter"])))
    if prec == "ref64":
        assert err_vs(r["pose_opt"], g["ref64_mc_pose"]) < 1e-8
        assert err_vs(r["cost_init"], g["ref64_mc_cost_init"]) < 1e-9
        assert err_vs(r["trans_mode"], g["ref64_mc_trans_mode"]) < 1e-7
        assert err_vs(r["trans_tril"], g["ref64_mc_trans_cov_tril"]) < 1e-6
        assert err_vs(r["rot_tril"], g["ref64_mc_rot_cov_tril"]) < 1e-6
        assert err_vs(r["samples"], g["ref64_mc_samples"]) < 1e-7
        assert err_vs(r["logw"], g["ref64_mc_logw"]) < 1e-7
    else:
        # calibrated: the reference run in fp32 vs itself in fp64 (same noise) is the floor
        floor_s = err_vs(g["ref32_mc_samples"], g["ref64_mc_samples"])
        floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
        assert err_vs(r["samples"], g["ref32_mc_samples"]) < max(1e-4, 3 * floor_s)
        assert err_vs(r["logw"], g["ref32_mc_logw"]) < max(1e-4, 3 * floor_w)


@pytest.mark.parametrize("prec", ["ref64", "ref32"])
def test_amis_4dof(prec):
    g = load_golden("mc4_basic")
    dtype = torch.float64 if prec == "ref64" else torch.float32
    x3d, x2d, w2d, cam, delta, pose_init = _setup(g, dtype)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    yaw_key = "yaw_samples64" if prec == "ref64" else "yaw_samples"
    noise = tuple(torch.from_numpy(g[k]).to(dtype) for k in ("noise_normal", "noise_chi2", yaw_key))
    pose, cov, _ = orc.lm_solve(x3d, x2d, w2d, cam, delta, pose_init, orc.LMParams(num_iter=int(g["lm_iter"])))
    r = orc.amis_4dof(x3d, x2d, w2d, cam, delta, pose, cov, noise, M, I)
    if prec == "ref64":
        assert err_vs(pose, g["ref64_mc_pose"]) < 1e-8
        assert err_vs(r["trans_mode"], g["ref64_mc_trans_mode"]) < 1e-7
        assert err_vs(r["trans_tril"], g["ref64_mc_trans_cov_tril"]) < 1e-6
        assert err_vs(r["rot_mode"], g["ref64_mc_rot_mode"][..., 0]) < 1e-7
        assert err_vs(r["rot_kappa"], g["ref64_mc_rot_kappa"][..., 0]) < 1e-6
        assert err_vs(r["samples"], g["ref64_mc_samples"]) < 1e-7
        assert err_vs(r["logw"], g["ref64_mc_logw"]) < 1e-7
    else:
        floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
        assert err_vs(r["logw"], g["ref32_mc_logw"]) < max(1e-4, 3 * floor_w)


def test_bessel_polynomial_matches_torch():
    from torch.distributions.von_mises import _log_modified_bessel_fn
    x = torch.logspace(-3, 3, 200, dtype=torch.float64)
    assert (orc.log_bessel_i0(x) - _log_modified_bessel_fn(x, order=0)).abs().max() < 1e-12
    exact = torch.special.i0e(x).log() + x
    assert (orc.log_bessel_i0(x) - exact).abs().max() < 5e-7      # the polynomial's own accuracy


def test_student_t_density_against_scipy():
    """The un-vendored pyro piece: multivariate Student-t log density (df=3, n=3)."""
    from scipy.stats import multivariate_t
    g = torch.Generator().manual_seed(5)
    A = torch.randn(3, 3, generator=g, dtype=torch.float64)
    cov = A @ A.T + 0.5 * torch.eye(3, dtype=torch.float64)
    L = torch.linalg.cholesky(cov)
    loc = torch.randn(3, generator=g, dtype=torch.float64)
    x = torch.randn(50, 3, generator=g, dtype=torch.float64) * 3
    ours = orc.mvt_logpdf(x, loc, L, df=3.0).numpy()
    ref = multivariate_t(loc=loc.numpy(), shape=cov.numpy(), df=3).logpdf(x.numpy())
    assert np.abs(ours - ref).max() < 1e-10


def test_acg_density_normalises():
    """ACG log density integrates to 1 over S^3 (Monte-Carlo with uniform sphere samples)."""
    g = torch.Generator().manual_seed(6)
    A = torch.randn(4, 4, generator=g, dtype=torch.float64)
    L = torch.linalg.cholesky(A @ A.T + 0.3 * torch.eye(4, dtype=torch.float64))
    u = torch.randn(400000, 4, generator=g, dtype=torch.float64)
    u = u / u.norm(dim=-1, keepdim=True)
    area = 2 * math.pi ** 2
    integral = orc.acg_logpdf(u, L).exp().mean().item() * area
    assert abs(integral - 1.0) < 2e-2


def test_student_t_sampler_moments():
    g = torch.Generator().manual_seed(7)
    n = 200000
    n3 = torch.randn(n, 3, generator=g, dtype=torch.float64)
    c2 = torch.randn(n, 5, generator=g, dtype=torch.float64).square().sum(-1)   # chi2(5): finite var
    L = torch.tensor([[1.0, 0, 0], [0.3, 0.8, 0], [-0.2, 0.1, 0.5]], dtype=torch.float64)
    loc = torch.tensor([1.0, -2.0, 0.5], dtype=torch.float64)
    x = orc.mvt_draw(n3, c2, loc, L, df=5.0)
    assert (x.mean(0) - loc).abs().max() < 0.02
    cov = torch.cov(x.T)
    assert (cov - (5.0 / 3.0) * (L @ L.T)).abs().max() < 0.06


def test_empty_batch_shapes():
    x3d, x2d, w2d = torch.zeros(0, 8, 3), torch.zeros(0, 8, 2), torch.zeros(0, 8, 2)
    cam = orc.Camera(torch.zeros(0, 3, 3))
    e = orc.evaluate(x3d, x2d, w2d, torch.zeros(0, 7), cam, 1.0, want_jac=True)
    assert e["cost"].shape == (0,) and e["jac"].shape == (0, 16, 6)
