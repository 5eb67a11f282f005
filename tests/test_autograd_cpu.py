"""The torch-composite pieces of the training path run on any device: checked here on CPU (float64) against
gradients recorded from the unmodified reference's autograd (oracle/make_golden.py, grads=True cases)."""
import pytest
import torch

from conftest import err_vs, golden_bounds, load_golden
from epropnp.autograd import gn_step_autograd
from epropnp.camera import PerspectiveCamera
from epropnp.cost_fun import AdaptiveHuberPnPCost
from epropnp.levenberg_marquardt import LMSolver


@pytest.mark.parametrize("name", ["mc6_basic", "mc6_bounds", "mc4_basic"])
def test_differentiable_gn_step_matches_reference_autograd(name):
    g = load_golden(name)
    d = torch.float64
    t = lambda k: torch.from_numpy(g[k]).to(d)
    x3d, x2d, w2d = (t(k).requires_grad_(True) for k in ("x3d", "x2d", "w2d"))
    lb, ub = golden_bounds(g, d)
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]), lb=lb, ub=ub)
    cost_fun = AdaptiveHuberPnPCost(relative_delta=float(g["relative_delta"]))
    cost_fun.set_param(x2d.detach(), w2d)                       # delta carries gradient to w2d (lib/train.py:175-177)
    assert cost_fun.delta.requires_grad
    solver = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]))
    pose = t("ref64_grad_pose_opt")
    plus = solver.pose_add(pose, gn_step_autograd(solver, x3d, x2d, w2d, pose, camera, cost_fun), camera)
    assert err_vs(plus.detach(), g["ref64_grad_pose_plus"]) < 1e-9
    grads = torch.autograd.grad((t("grad_c3") * plus).sum(), [x3d, x2d, w2d])
    for nm, gr in zip(("x3d", "x2d", "w2d"), grads):
        assert err_vs(gr, g[f"ref64_gradL2_{nm}"]) < 1e-7, nm
