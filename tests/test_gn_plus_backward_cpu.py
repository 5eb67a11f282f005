"""The native backward of pose_opt_plus (epnp_gn_plus_backward_f32: forward-mode duals per correspondence) against
torch autograd through the composite (PerspectiveCamera.project + HuberPnPCost.compute + linalg.solve + pose_add) in
float64 -- the kernel runs under the CPU SIMT emulator.  The composite itself is pinned to the reference's own
autograd by the gradient goldens (tests/test_autograd_cpu.py)."""
import numpy as np
import pytest
import torch

import simt_native
from conftest import err_vs, golden_bounds, load_golden
from epropnp import autograd as ag
from epropnp.camera import PerspectiveCamera
from epropnp.cost_fun import AdaptiveHuberPnPCost, HuberPnPCost
from epropnp.levenberg_marquardt import LMSolver


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


def _case(name, dev, dtype):
    g = load_golden(name)
    t = lambda k: torch.from_numpy(g[k]).to(device=dev, dtype=dtype)
    lb, ub = golden_bounds(g)
    if torch.is_tensor(lb):
        lb, ub = lb.to(device=dev, dtype=dtype), ub.to(device=dev, dtype=dtype)
    x3d, x2d, w2d = (t(k).clone().requires_grad_(True) for k in ("x3d", "x2d", "w2d"))
    camera = PerspectiveCamera(cam_mats=t("cam_mats"), z_min=float(g["z_min"]), lb=lb, ub=ub)
    if float(g["fixed_delta"]) >= 0:
        cost_fun = HuberPnPCost(delta=float(g["fixed_delta"]))
    else:
        cost_fun = AdaptiveHuberPnPCost(relative_delta=float(g["relative_delta"]))
        cost_fun.delta = t("delta").clone().requires_grad_(True)        # delta as a leaf: its own gradient is checked
    pose = torch.from_numpy(g["ref32_lm_pose"]).to(device=dev, dtype=dtype)
    return g, x3d, x2d, w2d, camera, cost_fun, pose


@pytest.mark.parametrize("name", ["lm6_basic", "lm6_bounds", "lm6_ragged", "lm4_basic", "lm6_scalar_bounds_fixed_delta",
                                  "mc6_basic", "mc4_basic"])
def test_native_backward_matches_composite(dev, monkeypatch, name):
    out = {}
    for mode, dtype in (("composite", torch.float64), ("native", torch.float32)):
        monkeypatch.setenv("EPNP_NATIVE_GN_STEP", "1" if mode == "native" else "0")
        g, x3d, x2d, w2d, camera, cost_fun, pose = _case(name, dev, dtype)
        solver = LMSolver(dof=int(g["dof"]), num_iter=int(g["lm_iter"]))
        plus = ag.pose_plus_autograd(solver, x3d, x2d, w2d, pose, camera, cost_fun)
        coef = torch.from_numpy(np.random.RandomState(3).randn(*plus.shape)).to(device=dev, dtype=dtype)
        (plus * coef).sum().backward()
        leaves = [x3d, x2d, w2d] + ([cost_fun.delta] if torch.is_tensor(cost_fun.delta) else [])
        out[mode] = (plus.detach(), [l.grad.detach() for l in leaves])
    assert err_vs(out["native"][0].cpu().numpy(), out["composite"][0].cpu().numpy()) < 1e-5
    for gn, gc, what in zip(out["native"][1], out["composite"][1], ("x3d", "x2d", "w2d", "delta")):
        assert torch.isfinite(gn).all(), what
        assert err_vs(gn.cpu().numpy(), gc.cpu().numpy()) < 1e-4, what    # measured 1e-6..2e-5, the same as torch's fp32 composite


def test_layer_uses_it_when_asked(dev, monkeypatch):
    from epropnp.epropnp import EProPnP6DoF
    monkeypatch.setenv("EPNP_NATIVE_GN_STEP", "1")
    g, x3d, x2d, w2d, camera, cost_fun, pose = _case("mc6_basic", dev, torch.float32)
    layer = EProPnP6DoF(mc_samples=64, num_iter=2, solver=LMSolver(dof=6, num_iter=5))
    r = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose, force_init_solve=False,
                                  with_pose_opt_plus=True, amis_seed=1)
    assert r[2].requires_grad and type(r[2].grad_fn).__name__ == "_PosePlusBackward"
    r[2].sum().backward()
    assert torch.isfinite(x3d.grad).all() and x3d.grad.abs().sum() > 0
