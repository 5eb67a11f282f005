#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference/epropnp, torch CPU) on seeded synthetic inputs.

    python oracle/make_golden.py            # needs /root/reference; run in the build container

The reference holds no tests / fixtures of its own, so these vectors are
what pins both the oracle restatement (oracle/pnp_oracle.py) and the CUDA path.  The only thing
that is not the reference's own code is `oracle/pyro_shim` (4 pyro names, see its docstring).

AMIS draws random numbers; to make its outputs comparable we record the *base noise* of every
draw (Student-t: standard normals + chi-square; ACG: standard normals) in call order and store it
next to the outputs.  Each case is also re-run in float64 with the recorded noise played back, so
a test can tell "differs from the reference" apart from "the reference's own fp32 rounding".
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "pyro_shim"))
sys.path.insert(0, "/root/reference")
sys.path.append(os.path.join(ROOT, "epro-pnp_b200"))     # after the reference: `epropnp` must resolve to /root/reference
warnings.filterwarnings("ignore")

import pyro.distributions as shim                      # noqa: E402  (the shim)
import epropnp.distributions as ref_distr              # noqa: E402  (reference)
from epropnp.camera import PerspectiveCamera           # noqa: E402
from epropnp.common import evaluate_pnp                # noqa: E402
from epropnp.cost_fun import AdaptiveHuberPnPCost, HuberPnPCost   # noqa: E402
from epropnp.epropnp import EProPnP4DoF, EProPnP6DoF   # noqa: E402
from epropnp.levenberg_marquardt import LMSolver       # noqa: E402
from epropnp_b200.synth import make_problem            # noqa: E402

assert "/root/reference" in os.path.abspath(ref_distr.__file__)


class NoiseTape:
    """Record / play back the base noise of every proposal draw, in call order."""

    def __init__(self):
        self.normal, self.chi2, self.rot = [], [], []
        self.play = None

    def install(self):
        tape = self

        def tap_normal(shape, dtype, device):
            if tape.play is not None:
                return tape.play["normal"].pop(0).to(dtype)
            x = torch.empty(shape, dtype=dtype, device=device).normal_()
            tape.normal.append(x.clone())
            return x

        def tap_chi2(df, sample_shape):
            if tape.play is not None:
                return tape.play["chi2"].pop(0).to(df.dtype)
            x = torch.distributions.Chi2(df).rsample(sample_shape)
            tape.chi2.append(x.clone())
            return x

        def tap_rot(shape, dtype, device):
            if tape.play is not None:
                return tape.play["rot"].pop(0).to(dtype)
            x = torch.empty(shape, dtype=dtype, device=device).normal_()
            tape.rot.append(x.clone())
            return x

        shim.draw_standard_normal = tap_normal
        shim.draw_chi2 = tap_chi2
        ref_distr._standard_normal = tap_rot   # name bound at epropnp/distributions.py:9, used :44

    def start_playback(self):
        self.play = dict(normal=[x.clone() for x in self.normal],
                         chi2=[x.clone() for x in self.chi2],
                         rot=[x.clone() for x in self.rot])


def build_camera_cost(p, dtype, z_min, bounds, relative_delta, fixed_delta):
    cam_mats = p["cam_mats"].to(dtype)
    if bounds == "tensor":
        lb = torch.stack((p["x2d"][..., 0].min(1).values + 8, p["x2d"][..., 1].min(1).values + 8), -1).to(dtype)
        ub = torch.stack((p["x2d"][..., 0].max(1).values - 8, p["x2d"][..., 1].max(1).values - 8), -1).to(dtype)
    elif bounds == "scalar":
        lb, ub = -100.5, 740.5
    else:
        lb = ub = None
    camera = PerspectiveCamera(cam_mats=cam_mats, z_min=z_min, lb=lb, ub=ub)
    if fixed_delta is not None:
        cost_fun = HuberPnPCost(delta=fixed_delta)
    else:
        cost_fun = AdaptiveHuberPnPCost(relative_delta=relative_delta)
        cost_fun.set_param(p["x2d"].to(dtype), p["w2d"].to(dtype))
    return camera, cost_fun, lb, ub


def to_np(x):
    if x is None:
        return None
    if torch.is_tensor(x):
        return x.detach().cpu().numpy()
    return np.asarray(x)


def run_case(name, B, N, dof=6, seed=0, lm_iter=10, fast_mode=False, mc=None, z_min=0.1,
             bounds=None, relative_delta=0.5, fixed_delta=None, outlier_frac=0.0, grid2d=False,
             behind=False, normalize=False, init_noise=(0.05, 3.0), eval_samples=5, grads=False):
    p = make_problem(B, N, seed=seed, dof=dof, outlier_frac=outlier_frac, grid2d=grid2d,
                     init_trans_noise=init_noise[0], init_rot_noise_deg=init_noise[1])
    if behind:
        # push a few points behind / onto the z_min plane to exercise the z clamp + clip_jac
        p["x3d"][:, :3] *= 14.0
    out = dict(B=B, N=N, dof=dof, lm_iter=lm_iter, fast_mode=int(fast_mode), z_min=z_min,
               relative_delta=relative_delta, normalize=int(normalize),
               fixed_delta=-1.0 if fixed_delta is None else fixed_delta,
               bounds=dict(none=0, scalar=1, tensor=2)["none" if bounds is None else bounds])
    for k in ("x3d", "x2d", "w2d", "cam_mats", "pose_init", "pose_gt"):
        out[k] = to_np(p[k])

    tape = NoiseTape()
    tape.install()
    for tag, dtype in (("ref32", torch.float32), ("ref64", torch.float64)):
        if tag == "ref64":
            tape.start_playback()
        x3d, x2d, w2d = (p[k].to(dtype) for k in ("x3d", "x2d", "w2d"))
        pose_init = p["pose_init"].to(dtype)
        camera, cost_fun, lb, ub = build_camera_cost(p, dtype, z_min, bounds, relative_delta, fixed_delta)
        if tag == "ref32":
            if torch.is_tensor(lb):
                out["lb"], out["ub"] = to_np(lb), to_np(ub)
            elif lb is not None:
                out["lb"], out["ub"] = np.float32(lb), np.float32(ub)
            out["delta"] = to_np(torch.as_tensor(cost_fun.delta, dtype=dtype).expand(B).clone())

        # --- evaluate_pnp known answers: at pose_init (with Jacobian) and at a few poses (cost)
        with torch.no_grad():
            res, cst, jac = evaluate_pnp(x3d, x2d, w2d, pose_init, camera, cost_fun,
                                         out_jacobian=True, out_residual=True, out_cost=True,
                                         clip_jac=not fast_mode)
            out[f"{tag}_eval_residual"], out[f"{tag}_eval_cost"], out[f"{tag}_eval_jac"] = \
                to_np(res), to_np(cst), to_np(jac)
            if tag == "ref32":
                g = torch.Generator().manual_seed(seed + 77)
                poses32 = pose_init[None].repeat(eval_samples, 1, 1).clone()
                poses32[..., :3] += 0.3 * torch.randn(eval_samples, B, 3, generator=g)
                if dof == 6:
                    q = poses32[..., 3:] + 0.2 * torch.randn(eval_samples, B, 4, generator=g)
                    poses32[..., 3:] = q / q.norm(dim=-1, keepdim=True)
                else:
                    poses32[..., 3] += 0.5 * torch.randn(eval_samples, B, generator=g)
                out["eval_poses"] = to_np(poses32)
            poses = poses32.to(dtype)       # identical (fp32-representable) poses for both runs
            out[f"{tag}_eval_cost_multi"] = to_np(
                evaluate_pnp(x3d, x2d, w2d, poses, camera, cost_fun, out_cost=True)[1])

        # --- LM / GN solve
        solver = LMSolver(dof=dof, num_iter=lm_iter, normalize=False)
        with torch.no_grad():
            pose_opt, pose_cov, cost, pose_plus = solver(
                x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init, with_pose_cov=True,
                with_cost=True, with_pose_opt_plus=True, fast_mode=fast_mode)
        out[f"{tag}_lm_pose"], out[f"{tag}_lm_cov"], out[f"{tag}_lm_cost"], out[f"{tag}_lm_pose_plus"] = \
            to_np(pose_opt), to_np(pose_cov), to_np(cost), to_np(pose_plus)
        if normalize:
            solver_n = LMSolver(dof=dof, num_iter=lm_iter, normalize=True)
            with torch.no_grad():
                pose_n = solver_n(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init,
                                  with_cost=True, fast_mode=fast_mode)
            out[f"{tag}_lmnorm_pose"], out[f"{tag}_lmnorm_cost"] = to_np(pose_n[0]), to_np(pose_n[2])

        # --- AMIS
        if mc is not None:
            M, I = mc
            cls = EProPnP6DoF if dof == 6 else EProPnP4DoF
            layer = cls(mc_samples=M, num_iter=I, normalize=normalize,
                        solver=LMSolver(dof=dof, num_iter=lm_iter))
            stash = {}
            orig_alloc = layer.allocate_buffer

            def alloc(*a, _o=orig_alloc, _s=stash, **kw):
                bufs = _o(*a, **kw)
                _s["bufs"] = bufs
                return bufs
            layer.allocate_buffer = alloc
            if dof == 4:
                np.random.seed(seed + 5)       # VonMisesUniformMix samples with numpy (distributions.py:64-72)
            with torch.no_grad():
                r = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init,
                                              force_init_solve=False, with_cost=True, fast_mode=fast_mode)
            pose_opt, cost, _, samples, logw, cost_init = r
            out[f"{tag}_mc_pose"], out[f"{tag}_mc_cost"] = to_np(pose_opt), to_np(cost)
            out[f"{tag}_mc_samples"], out[f"{tag}_mc_logw"], out[f"{tag}_mc_cost_init"] = \
                to_np(samples), to_np(logw), to_np(cost_init)
            names = ("trans_mode", "trans_cov_tril", "rot_cov_tril") if dof == 6 else \
                    ("trans_mode", "trans_cov_tril", "rot_mode", "rot_kappa")
            for nm, bf in zip(names, stash["bufs"]):
                out[f"{tag}_mc_{nm}"] = to_np(bf)
            if dof == 4:
                # 4DoF draws yaw on the host with numpy (distributions.py:61-72): no base noise to replay, so
                # keep the drawn yaw samples of EACH run (they depend on that run's mode / kappa)
                out["yaw_samples" if tag == "ref32" else "yaw_samples64"] = to_np(samples[..., 3].reshape(I, M // I, B))
            if tag == "ref32":
                out["mc_samples_total"], out["mc_iters"] = M, I
                S = M // I
                # (I, S, B, .) in draw order
                out["noise_normal"] = to_np(torch.stack(tape.normal).reshape(I, S, B, 3))
                out["noise_chi2"] = to_np(torch.stack(tape.chi2).reshape(I, S, B))
                if dof == 6:
                    out["noise_rot"] = to_np(torch.stack(tape.rot).reshape(I, S, B, 4))

    if grads and mc is not None:
        # Gradients of the reference's differentiable outputs (epropnp.py:108-113), float64, recorded noise played
        # back: L1 = <c1, logweights> + <c2, cost_init>, L2 = <c3, pose_opt_plus>; delta depends on w2d exactly
        # as the reference's training loop sets it (EPro-PnP-6DoF/lib/train.py:175-177).
        tape.start_playback()
        d = torch.float64
        x3d = p["x3d"].to(d).requires_grad_(True)
        x2d = p["x2d"].to(d).requires_grad_(True)
        w2d = p["w2d"].to(d).requires_grad_(True)
        camera, _, _, _ = build_camera_cost(p, d, z_min, bounds, relative_delta, fixed_delta)
        cost_fun = AdaptiveHuberPnPCost(relative_delta=relative_delta)
        cost_fun.set_param(x2d.detach(), w2d)
        M, I = mc
        cls = EProPnP6DoF if dof == 6 else EProPnP4DoF
        layer = cls(mc_samples=M, num_iter=I, solver=LMSolver(dof=dof, num_iter=lm_iter))
        if dof == 4:
            np.random.seed(seed + 5)
        pose_opt, _, pose_plus, samples, logw, cost_init = layer.monte_carlo_forward(
            x3d, x2d, w2d, camera, cost_fun, pose_init=p["pose_init"].to(d), force_init_solve=False,
            with_pose_opt_plus=True)
        g = torch.Generator().manual_seed(seed + 99)
        c1 = 0.01 * torch.randn(logw.shape, generator=g, dtype=d)
        c2 = torch.randn(cost_init.shape, generator=g, dtype=d)
        c3 = torch.randn(pose_plus.shape, generator=g, dtype=d)
        g1 = torch.autograd.grad((c1 * logw).sum() + (c2 * cost_init).sum(), [x3d, x2d, w2d], retain_graph=True)
        g2 = torch.autograd.grad((c3 * pose_plus).sum(), [x3d, x2d, w2d])
        out["grad_c1"], out["grad_c2"], out["grad_c3"] = to_np(c1), to_np(c2), to_np(c3)
        for nm, a, b in zip(("x3d", "x2d", "w2d"), g1, g2):
            out[f"ref64_gradL1_{nm}"], out[f"ref64_gradL2_{nm}"] = to_np(a), to_np(b)
        out["ref64_grad_pose_opt"], out["ref64_grad_pose_plus"] = to_np(pose_opt), to_np(pose_plus)
        out["ref64_grad_logw"], out["ref64_grad_cost_init"] = to_np(logw), to_np(cost_init)

    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **{k: v for k, v in out.items() if v is not None})
    d32, d64 = out["ref32_lm_pose"], out["ref64_lm_pose"]
    msg = f"{name}: B={B} N={N} dof={dof}  |lm_pose32-64|max={np.abs(d32 - d64).max():.2e}"
    if mc is not None:
        a, b = out["ref32_mc_logw"], out["ref64_mc_logw"]
        msg += f"  |logw32-64|max={np.abs(a - b).max():.2e} (|logw|~{np.abs(b).mean():.1f})"
    print(msg, f" -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(1234)
    torch.set_num_threads(4)
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    # 6DoF Levenberg-Marquardt
    run_case("lm6_basic", B=6, N=64, seed=11)
    run_case("lm6_ragged", B=5, N=51, seed=12, outlier_frac=0.1)
    run_case("lm6_bounds", B=5, N=100, seed=13, bounds="tensor", behind=True, outlier_frac=0.1)
    run_case("lm6_scalar_bounds_fixed_delta", B=4, N=128, seed=14, bounds="scalar", fixed_delta=1.0,
             init_noise=(0.15, 8.0))
    run_case("gn6_fast_grid", B=4, N=256, seed=15, lm_iter=3, fast_mode=True, grid2d=True,
             relative_delta=0.1, z_min=0.01)
    run_case("lm6_normalize", B=4, N=64, seed=16, normalize=True)
    # 6DoF AMIS
    run_case("mc6_basic", B=4, N=64, seed=21, mc=(512, 4), grads=True)
    run_case("mc6_small", B=3, N=52, seed=22, mc=(64, 4), outlier_frac=0.05)
    run_case("mc6_bounds", B=3, N=100, seed=23, mc=(256, 2), bounds="tensor", grads=True)
    run_case("mc6_n512", B=2, N=512, seed=24, mc=(512, 4))
    # 4DoF
    run_case("lm4_basic", B=6, N=64, seed=31, dof=4)
    run_case("gn4_fast", B=4, N=128, seed=32, dof=4, lm_iter=5, fast_mode=True)
    run_case("mc4_basic", B=4, N=64, seed=33, dof=4, mc=(512, 4), grads=True)


if __name__ == "__main__":
    main()
