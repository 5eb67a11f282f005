"""TEST-ONLY stand-in for `epropnp_b200.native`, backed by the CPU oracle.

Lets the CPU suite drive the *host logic* of the drop-in package (epropnp.*: argument plumbing, shapes, normalise /
denormalise, RSLM bookkeeping, the autograd bridge) without a GPU: `install(monkeypatch)` swaps the handful of
tensor-level entry points for oracle calls.  The product never imports this file; on a real run every one of these
functions is a C-ABI call into libepropnp_b200.so and CPU tensors are refused.
"""
import math

import torch

from epropnp_b200 import native
from oracle import pnp_oracle as orc


class FakeProblem:
    """Same normalisation of inputs as native.Problem, but keeps dtype and accepts CPU tensors."""

    def __init__(self, x3d, x2d, w2d, cam_mats, lb, ub, delta):
        self.B, self.N = x3d.shape[0], x3d.shape[1]
        self.device = x3d.device
        dt = x3d.dtype
        self.x3d, self.x2d, self.w2d = x3d.detach(), x2d.detach(), w2d.detach()
        self.cam = cam_mats.detach().to(dt).expand(self.B, 3, 3)
        self.lb = None if lb is None or ub is None else (lb.detach().to(dt) if torch.is_tensor(lb) else float(lb))
        self.ub = None if lb is None or ub is None else (ub.detach().to(dt) if torch.is_tensor(ub) else float(ub))
        self.delta = delta.detach().to(dt).expand(self.B) if torch.is_tensor(delta) else torch.full((self.B,), float(delta), dtype=dt)

    def camera(self, z_min):
        return orc.Camera(self.cam, z_min, self.lb, self.ub)

    def empty(self, *shape):
        return torch.empty(shape, dtype=self.x3d.dtype)


def _lm_params(p):
    return orc.LMParams(num_iter=p.lm_iter, min_lm_diagonal=p.min_lm_diagonal, max_lm_diagonal=p.max_lm_diagonal,
                        min_relative_decrease=p.min_relative_decrease, initial_trust_region_radius=p.initial_radius,
                        max_trust_region_radius=p.max_radius, eps=p.eps)


def adaptive_delta(x2d, w2d, relative_delta):
    return orc.adaptive_delta(x2d.detach(), w2d.detach(), relative_delta)


def evaluate_cost(prob, poses, dof, z_min):
    return orc.evaluate(prob.x3d, prob.x2d, prob.w2d, poses.detach().to(prob.x3d.dtype), prob.camera(z_min), prob.delta)["cost"]


def evaluate_full(prob, pose, dof, z_min, huber_eps, clip_jac, want_residual, want_jac, want_cost):
    e = orc.evaluate(prob.x3d, prob.x2d, prob.w2d, pose.detach().to(prob.x3d.dtype), prob.camera(z_min), prob.delta,
                     want_jac=True, clip_jac=bool(clip_jac), eps_huber=huber_eps)
    return (e["residual"] if want_residual else None, e["cost"] if want_cost else None, e["jac"] if want_jac else None)


def lm_solve(prob, pose_init, params, want_cov=False, want_cost=False, want_plus=False, want_cost_init=False):
    cam = prob.camera(params.z_min)
    pi = pose_init.detach().to(prob.x3d.dtype)
    pose, cov, cost = orc.lm_solve(prob.x3d, prob.x2d, prob.w2d, cam, prob.delta, pi, _lm_params(params),
                                   fast_mode=bool(params.fast_mode))
    plus = None
    if want_plus:
        plus = orc.pose_add(pose, orc.gn_step(prob.x3d, prob.x2d, prob.w2d, cam, prob.delta, pose, params.eps))
    ci = orc.evaluate(prob.x3d, prob.x2d, prob.w2d, pi, cam, prob.delta)["cost"] if want_cost_init else None
    return dict(pose_opt=pose, pose_cov=cov if want_cov else None, cost=cost if want_cost else None,
                pose_opt_plus=plus, cost_init=ci)


def _noise_isb(noise, B, M, I, dof, dtype, seed):
    S = M // I
    if noise is None:
        g = torch.Generator().manual_seed(int(seed) % (2 ** 31))
        n3 = torch.randn(B, M, 3, generator=g, dtype=torch.float64)
        c2 = torch.randn(B, M, 3, generator=g, dtype=torch.float64).square().sum(-1)
        r = torch.randn(B, M, 4, generator=g, dtype=torch.float64) if dof == 6 else \
            (torch.rand(B, M, generator=g, dtype=torch.float64) * 2 - 1) * math.pi
        noise = (n3, c2, r)
    n3, c2, r = (t.detach().to(dtype) for t in noise)
    out = (n3.reshape(B, I, S, 3).permute(1, 2, 0, 3), c2.reshape(B, I, S).permute(1, 2, 0))
    return out + ((r.reshape(B, I, S, 4).permute(1, 2, 0, 3),) if dof == 6 else (r.reshape(B, I, S).permute(1, 2, 0),))


def lm_amis_fused(prob, pose_init, params, noise=None, seed=0, obj_offset=0, want_cost=False, want_plus=False,
                  want_cost_init=True, want_proposals=False, want_cov=True):
    lm = lm_solve(prob, pose_init, params, want_cov=True, want_cost=True, want_plus=want_plus, want_cost_init=want_cost_init)
    cam = prob.camera(params.z_min)
    M, I, dof = params.mc_samples, params.mc_iter, params.dof
    nz = _noise_isb(noise, prob.B, M, I, dof, prob.x3d.dtype, seed)
    if dof == 6:
        r = orc.amis_6dof(prob.x3d, prob.x2d, prob.w2d, cam, prob.delta, lm["pose_opt"], lm["pose_cov"], nz, M, I,
                          params.amis_eps, params.acg_mle_iter, params.acg_dispersion)
    else:
        r = orc.amis_4dof(prob.x3d, prob.x2d, prob.w2d, cam, prob.delta, lm["pose_opt"], lm["pose_cov"], nz, M, I,
                          params.amis_eps)
    return dict(pose_opt=lm["pose_opt"], pose_cov=lm["pose_cov"] if want_cov else None,
                cost=lm["cost"] if want_cost else None, pose_opt_plus=lm["pose_opt_plus"], cost_init=lm["cost_init"],
                pose_samples=r["samples"].transpose(0, 1).contiguous(), logw=r["logw"].transpose(0, 1).contiguous(),
                proposals=None)


def rslm(prob, inds, start, params, want_all=False):
    r = orc.rslm_solve(prob.x3d, prob.x2d, prob.w2d, prob.camera(params.z_min), prob.delta, inds, start.detach().to(prob.x3d.dtype),
                       _lm_params(params), fast_mode=bool(params.fast_mode))
    return dict(pose=r["pose"], cost=r["cost"], pose_all=r["hyp_pose"] if want_all else None,
                cost_all=r["hyp_cost"] if want_all else None)


def rslm_draw(x3d, x2d, w2d, cam_mats, P, n, dof, eps=1e-5, seed=0, obj_offset=0, t_init=None, want_t=False):
    """The reference's own draws (levenberg_marquardt.py:306-324) from a torch generator seeded with `seed`."""
    g = torch.Generator().manual_seed(int(seed) % (2 ** 63))
    B, N = w2d.shape[0], w2d.shape[1]
    if t_init is None:
        t_init = orc.center_based_init(x2d.detach(), x3d.detach(), orc.Camera(cam_mats), dof)
    rows = w2d.detach().mean(dim=-1).unsqueeze(0).expand(P, B, N).reshape(P * B, N)
    inds = torch.multinomial(rows.double(), n, generator=g).reshape(P, B, n).to(torch.int32)
    start = t_init.new_empty((P, B, 7 if dof == 6 else 4))
    start[..., :3] = t_init
    if dof == 4:
        start[..., 3] = torch.rand((P, B), generator=g, dtype=t_init.dtype) * (2 * math.pi)
    else:
        q = torch.randn((P, B, 4), generator=g, dtype=t_init.dtype)
        qn = q.norm(dim=-1, keepdim=True)
        start[..., 3:] = torch.where(qn < eps, q.new_tensor([1., 0., 0., 0.]), q / qn)
    return (inds, start, t_init) if want_t else (inds, start)


def cost_backward(prob, dof, z_min, poses_a, grad_a, poses_b=None, grad_b=None, want=(True, True, True, True)):
    """Reference semantics by construction: torch autograd through the oracle's evaluate."""
    with torch.enable_grad():            # we are called from inside Function.backward, where grad mode is off
        x3d, x2d, w2d = (t.clone().requires_grad_(True) for t in (prob.x3d, prob.x2d, prob.w2d))
        delta = prob.delta.clone().requires_grad_(True)
        poses, grads = poses_a.detach(), grad_a.detach()
        if poses_b is not None:
            poses, grads = torch.cat((poses, poses_b.detach()), 1), torch.cat((grads, grad_b.detach()), 1)
        cost = orc.evaluate(x3d, x2d, w2d, poses.transpose(0, 1).to(x3d.dtype), prob.camera(z_min), delta)["cost"]   # (P, B)
        (cost * grads.transpose(0, 1).to(cost.dtype)).sum().backward()
    full = (x3d.grad, x2d.grad, w2d.grad, delta.grad)
    return tuple(g if w else None for g, w in zip(full, want))


def install(monkeypatch):
    # the derivative-regularisation branch: the torch composite (exact in float64) instead of the fp32 native kernel
    monkeypatch.setenv("EPNP_NATIVE_GN_STEP", "0")
    for name, fn in (("Problem", FakeProblem), ("adaptive_delta", adaptive_delta), ("evaluate_cost", evaluate_cost),
                     ("evaluate_full", evaluate_full), ("lm_solve", lm_solve), ("lm_amis_fused", lm_amis_fused), ("rslm", rslm), ("rslm_draw", rslm_draw),
                     ("cost_backward", cost_backward)):
        monkeypatch.setattr(native, name, fn)
