"""LMSolver._lm_iter and camera.project_a / project_b: names of the reference's surface that the
kernels make redundant -- `solve` runs its iterations inside lm_warp_kernel -- kept as stand-alone torch restatements.
Pinned here in float64 on the goldens of the UNMODIFIED reference: driving `_lm_iter` the way the reference's `solve` does
(levenberg_marquardt.py:136-175) must land on the reference's LM result."""
import pytest
import torch

from conftest import err_vs, golden_bounds, golden_names, load_golden
from epropnp.camera import PerspectiveCamera, project_a, project_b
from epropnp.levenberg_marquardt import LMSolver
from oracle import pnp_oracle as orc

# the trust-region iteration is the non-fast, un-normalised path
LM_CASES = [n for n in golden_names() if n.startswith("lm")
            and not bool(load_golden(n)["fast_mode"]) and not int(load_golden(n)["normalize"])]


@pytest.mark.parametrize("name", LM_CASES)
def test_lm_iter_driven_like_the_reference_solve(name):
    g = load_golden(name)
    d = torch.float64
    t = lambda k: torch.from_numpy(g[k]).to(d)
    lb, ub = golden_bounds(g, d)
    cam = orc.Camera(t("cam_mats"), float(g["z_min"]), lb, ub)
    delta = orc.adaptive_delta(t("x2d"), t("w2d"), float(g["relative_delta"])) if float(g["fixed_delta"]) < 0 \
        else float(g["fixed_delta"])
    x3d, x2d, w2d = t("x3d"), t("x2d"), t("w2d")
    B, N = x2d.shape[:2]
    dof = 4 if t("pose_init").shape[-1] == 4 else 6
    solver = LMSolver(dof=dof, num_iter=int(g["lm_iter"]))

    def evaluate_fun(pose, out_jacobian=None, out_residual=None, out_cost=None):     # = partial(evaluate_pnp, ...)
        e = orc.evaluate(x3d, x2d, w2d, pose, cam, delta, want_jac=True, clip_jac=True)
        out_jacobian.copy_(e["jac"]); out_residual.copy_(e["residual"]); out_cost.copy_(e["cost"])

    pose = t("pose_init").clone()
    jac, jac_new = torch.empty(B, 2 * N, dof, dtype=d), torch.empty(B, 2 * N, dof, dtype=d)
    res, res_new = torch.empty(B, 2 * N, dtype=d), torch.empty(B, 2 * N, dtype=d)
    cost, cost_new = torch.empty(B, dtype=d), torch.empty(B, dtype=d)
    evaluate_fun(pose=pose, out_jacobian=jac_new, out_residual=res_new, out_cost=cost_new)
    took = torch.ones(B, dtype=torch.bool)
    radius = torch.full((B,), solver.initial_trust_region_radius, dtype=d)
    shrink = torch.full((B,), 2.0, dtype=d)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]))
    for _ in range(solver.num_iter):
        solver._lm_iter(pose, jac, res, cost, jac_new, res_new, cost_new, took, radius, shrink, evaluate_fun, camera)
    assert err_vs(pose, g["ref64_lm_pose"]) < 1e-8
    final_cost = torch.where(took, cost_new, cost)
    assert err_vs(final_cost, g["ref64_lm_cost"]) < 1e-8


@pytest.mark.parametrize("dof", [6, 4])
def test_project_a_and_b_agree_with_the_camera(dof):
    from epropnp_b200.synth import make_problem
    pc = {k: v.double() for k, v in make_problem(3, 20, seed=5, dof=dof).items()}
    poses = pc["pose_gt"][None].repeat(4, 1, 1) + 0.01 * torch.randn(4, 3, pc["pose_gt"].shape[-1], dtype=torch.float64)
    ua, rot, za = project_a(pc["x3d"], poses, pc["cam_mats"], 0.1)
    ub, zb = project_b(pc["x3d"], poses, pc["cam_mats"], 0.1)
    u, _ = PerspectiveCamera(cam_mats=pc["cam_mats"], z_min=0.1).project(pc["x3d"], poses)
    assert ua.shape == (4, 3, 20, 2) and rot.shape == (4, 3, 20, 3) and za.shape == zb.shape == (4, 3, 20, 1)
    assert torch.allclose(ua, u, atol=1e-12) and torch.allclose(ub, u, atol=1e-9) and torch.allclose(za, zb, atol=1e-12)
    want = orc.evaluate(pc["x3d"], pc["x2d"], pc["w2d"], poses, orc.Camera(pc["cam_mats"], 0.1), 1.0)
    assert want["cost"].shape == (4, 3)
