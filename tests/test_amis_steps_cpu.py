"""The AMIS steps the reference exposes as methods of the layer -- allocate_buffer, initial_fit, gen_new_distr,
gen_old_distr, estimate_params (epropnp.py:65-82, 209-260, 282-342) -- and the translation proposal it takes from pyro.
In the drop-in they are stand-alone torch restatements (the loop itself runs inside the kernel).  Pinned here by replaying
the whole loop of the UNMODIFIED reference on its golden vectors in float64: the golden samples go through the
methods' densities and refits (costs from the oracle), and every proposal's parameters and the final log-weights must
come out as the reference produced them.  CPU tensors: these methods are plain torch, not the native path."""
import math

import numpy as np
import pytest
import torch

from conftest import err_vs, golden_bounds, golden_names, load_golden
from epropnp.camera import PerspectiveCamera
from epropnp.distributions import MultivariateStudentT
from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF
from oracle import pnp_oracle as orc


def _replay(g, dof):
    d = torch.float64
    t = lambda k: torch.from_numpy(g[k]).to(d)
    I, M = int(g["mc_iters"]), int(g["mc_samples_total"])
    S = M // I
    layer = (EProPnP6DoF if dof == 6 else EProPnP4DoF)(mc_samples=M, num_iter=I)
    lb, ub = golden_bounds(g, d)
    cam = orc.Camera(t("cam_mats"), float(g["z_min"]), lb, ub)
    delta = orc.adaptive_delta(t("x2d"), t("w2d"), float(g["relative_delta"])) if float(g["fixed_delta"]) < 0 \
        else float(g["fixed_delta"])
    pose_opt, pose_cov = t("ref64_lm_pose"), t("ref64_lm_cov")
    B = pose_opt.shape[0]
    samples = t("ref64_mc_samples").reshape(I, S, B, -1)
    params = layer.allocate_buffer(B, dtype=d)
    layer.initial_fit(pose_opt, pose_cov, PerspectiveCamera(cam_mats=t("cam_mats")), *params)
    logprobs = torch.empty(I, I, S, B, dtype=d)
    cost = torch.empty(I, S, B, dtype=d)
    for i in range(I):
        new_t, new_r = layer.gen_new_distr(i, *params)
        cost[i] = orc.evaluate(t("x3d"), t("x2d"), t("w2d"), samples[i], cam, delta)["cost"]
        # exactly the reference's calls (epropnp.py:152-163): the yaw keeps its trailing axis of 1, flatten(2) drops it
        logprobs[i, :i + 1] = new_t.log_prob(samples[:i + 1, ..., :3]) + new_r.log_prob(samples[:i + 1, ..., 3:]).flatten(2)
        if i > 0:
            old_t, old_r = layer.gen_old_distr(i, *params)
            logprobs[:i, i] = old_t.log_prob(samples[i, ..., :3]) + old_r.log_prob(samples[i, ..., 3:]).flatten(2)
        logw = -cost[:i + 1] - (torch.logsumexp(logprobs[:i + 1, :i + 1], dim=0) - math.log(i + 1))
        if i + 1 < I:
            layer.estimate_params(i, samples[:i + 1].reshape((i + 1) * S, B, -1), logw.reshape((i + 1) * S, B), *params)
    return params, logw.reshape(M, B)


@pytest.mark.parametrize("name", golden_names("mc6"))
def test_6dof_steps_reproduce_the_reference_loop(name):
    g = load_golden(name)
    (trans_mode, trans_tril, rot_tril), logw = _replay(g, 6)
    assert err_vs(trans_mode, g["ref64_mc_trans_mode"]) < 1e-8
    assert err_vs(trans_tril, g["ref64_mc_trans_cov_tril"]) < 1e-7
    assert err_vs(rot_tril, g["ref64_mc_rot_cov_tril"]) < 1e-7
    assert err_vs(logw, g["ref64_mc_logw"]) < 1e-8


@pytest.mark.parametrize("name", golden_names("mc4"))
def test_4dof_steps_reproduce_the_reference_loop(name):
    g = load_golden(name)
    (trans_mode, trans_tril, rot_mode, rot_kappa), logw = _replay(g, 4)
    assert err_vs(trans_mode, g["ref64_mc_trans_mode"]) < 1e-8
    assert err_vs(trans_tril, g["ref64_mc_trans_cov_tril"]) < 1e-7
    assert err_vs(rot_mode, g["ref64_mc_rot_mode"]) < 1e-8
    assert err_vs(rot_kappa, g["ref64_mc_rot_kappa"]) < 1e-7
    assert err_vs(logw, g["ref64_mc_logw"]) < 1e-8


def test_student_t_density_and_sampler():
    """Against scipy's multivariate t (the reference's class comes from pyro, which is not installed here), and the
    sampler's covariance: df / (df - 2) * L L^T for df > 2."""
    from scipy.stats import multivariate_t
    g = torch.Generator().manual_seed(0)
    A = torch.randn(3, 3, generator=g, dtype=torch.float64)
    L = torch.linalg.cholesky(A @ A.T + 0.5 * torch.eye(3, dtype=torch.float64))
    loc = torch.tensor([0.3, -1.0, 2.0], dtype=torch.float64)
    x = torch.randn(50, 3, generator=g, dtype=torch.float64) * 2
    for df in (3, 5.5):
        ours = MultivariateStudentT(df, loc, L).log_prob(x)
        want = multivariate_t(loc=loc.numpy(), shape=(L @ L.T).numpy(), df=df).logpdf(x.numpy())
        assert np.abs(ours.numpy() - want).max() < 1e-10
    torch.manual_seed(1)
    draws = MultivariateStudentT(6.0, loc, L).rsample((200000,))
    cov = torch.cov(draws.T)
    assert (draws.mean(0) - loc).abs().max() < 0.03
    assert (cov - 1.5 * (L @ L.T)).abs().max() < 0.05 * (L @ L.T).abs().max()
    # batch shapes broadcast like the reference's mixture call (iter, 1, B) against samples (S, B, 3)
    mix = MultivariateStudentT(3, loc.expand(2, 1, 4, 3), L.expand(2, 1, 4, 3, 3))
    assert mix.log_prob(torch.zeros(7, 4, 3, dtype=torch.float64)).shape == (2, 7, 4)


def test_a_subclass_that_overrides_a_step_is_refused():
    class Custom(EProPnP6DoF):
        def estimate_params(self, *args, **kwargs):
            pass

    with pytest.raises(NotImplementedError, match="estimate_params"):
        Custom(mc_samples=8, num_iter=2).monte_carlo_forward(None, None, None, None, None)
    # a subclass that leaves the steps alone is fine (it fails later, on the missing solver, not on the guard)
    class Plain(EProPnP6DoF):
        pass
    Plain(mc_samples=8, num_iter=2)._refuse_overridden_steps()
    EProPnP4DoF(mc_samples=8, num_iter=2)._refuse_overridden_steps()
