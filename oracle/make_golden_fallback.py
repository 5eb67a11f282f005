#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the `cholesky_wrapper` fallback (epropnp.py:16-33: a covariance that is
not positive definite gets diag(default_diag), or the identity) from the UNMODIFIED reference layer.

    python oracle/make_golden_fallback.py        # needs /root/reference; writes tests/golden/fallback/*.npz

The reference's `monte_carlo_forward` (epropnp.py:87-196) is run end to end; only its solver is replaced by a stub that
hands back a crafted `pose_cov` (exactly representable entries), so that for SOME objects of the batch
    * the translation block is indefinite          -> L_t = I (6DoF, default_diag None) / diag(1, 1, 4) (4DoF)
    * the rotation block makes rot_cov indefinite   -> det^(1/4) is NaN, rot_cov_tril = I (6DoF)
while the other objects keep a healthy covariance (the fallback is per object, :24-30).  AMIS then draws and weighs its
samples from those first proposals (ONE iteration: a unit-covariance proposal around a millimetre-wide posterior puts all
the weight on one sample, and whether the refit covariance of that degenerate set is "positive definite" is decided by
the last bit of the summation order -- not a comparison any implementation, including the reference in another
precision, can pass); base noise is taped and replayed in float64 exactly as in make_golden.py.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg          # noqa: E402  (installs the reference + shim on sys.path, NoiseTape, helpers)
from make_golden import (EProPnP4DoF, EProPnP6DoF, LMSolver, NoiseTape, build_camera_cost, make_problem, to_np)  # noqa: E402

ROOT = mg.ROOT


class StubSolver(torch.nn.Module):
    """Returns the given (pose_opt, pose_cov) the way LMSolver.forward does (levenberg_marquardt.py:55-78)."""

    def __init__(self, pose_opt, pose_cov):
        super().__init__()
        self.pose_opt, self.pose_cov = pose_opt, pose_cov

    def forward(self, x3d, x2d, w2d, camera, cost_fun, **kwargs):
        return self.pose_opt.clone(), self.pose_cov.clone(), None, None


def crafted_cov(cov, dof, kinds):
    """cov (B, dof, dof) healthy covariances; kinds[b] in {"ok", "trans", "rot"}"""
    cov = cov.clone()
    for b, kind in enumerate(kinds):
        if kind == "trans":                     # indefinite translation block: diag(1, -1, 1) / 64
            cov[b, :3, :] = 0; cov[b, :, :3] = 0
            cov[b, 0, 0], cov[b, 1, 1], cov[b, 2, 2] = 1 / 64, -1 / 64, 1 / 64
        elif kind == "rot" and dof == 6:        # negative-definite rotation block: -I / 2 (and no cross terms)
            cov[b, 3:, :] = 0; cov[b, :, 3:] = 0
            cov[b, 3, 3] = cov[b, 4, 4] = cov[b, 5, 5] = -0.5
    return cov


def run_case(name, dof, kinds, N=64, M=128, I=1, seed=61):
    B = len(kinds)
    p = make_problem(B, N, seed=seed, dof=dof)
    out = dict(B=B, N=N, dof=dof, lm_iter=10, fast_mode=0, z_min=0.1, relative_delta=0.5, normalize=0, fixed_delta=-1.0,
               bounds=0, mc_samples_total=M, mc_iters=I, kinds=np.array(kinds))
    for k in ("x3d", "x2d", "w2d", "cam_mats", "pose_init"):
        out[k] = to_np(p[k])
    tape = NoiseTape()
    tape.install()
    for tag, dtype in (("ref32", torch.float32), ("ref64", torch.float64)):
        if tag == "ref64":
            tape.start_playback()
        x3d, x2d, w2d = (p[k].to(dtype) for k in ("x3d", "x2d", "w2d"))
        camera, cost_fun, _, _ = build_camera_cost(p, dtype, 0.1, None, 0.5, None)
        if tag == "ref32":
            out["delta"] = to_np(cost_fun.delta.reshape(B))
            # a healthy solution to start from: the reference's own LM result in float32
            with torch.no_grad():
                pose_opt, pose_cov, _ = LMSolver(dof=dof, num_iter=10).solve(x3d, x2d, w2d, camera, cost_fun, pose_init=p["pose_init"],
                                                                           with_pose_cov=True)
            pose_opt32, pose_cov32 = pose_opt, crafted_cov(pose_cov, dof, kinds)
            out["pose_opt_in"], out["pose_cov_in"] = to_np(pose_opt32), to_np(pose_cov32)
        cls = EProPnP6DoF if dof == 6 else EProPnP4DoF
        layer = cls(mc_samples=M, num_iter=I, solver=StubSolver(pose_opt32.to(dtype), pose_cov32.to(dtype)))
        stash = {}
        orig_alloc = layer.allocate_buffer

        def alloc(*a, _o=orig_alloc, _s=stash, **kw):
            _s["bufs"] = _o(*a, **kw)
            return _s["bufs"]
        layer.allocate_buffer = alloc
        if dof == 4:
            np.random.seed(seed + 5)
        with torch.no_grad():
            _, _, _, samples, logw, _ = layer.monte_carlo_forward(x3d, x2d, w2d, camera, cost_fun, pose_init=p["pose_init"].to(dtype),
                                                                  force_init_solve=False)
        out[f"{tag}_mc_samples"], out[f"{tag}_mc_logw"] = to_np(samples), to_np(logw)
        names = ("trans_mode", "trans_cov_tril", "rot_cov_tril") if dof == 6 else ("trans_mode", "trans_cov_tril", "rot_mode", "rot_kappa")
        for nm, bf in zip(names, stash["bufs"]):
            out[f"{tag}_mc_{nm}"] = to_np(bf)
        if dof == 4:
            out["yaw_samples" if tag == "ref32" else "yaw_samples64"] = to_np(samples[..., 3].reshape(I, M // I, B))
        if tag == "ref32":
            S = M // I
            out["noise_normal"] = to_np(torch.stack(tape.normal).reshape(I, S, B, 3))
            out["noise_chi2"] = to_np(torch.stack(tape.chi2).reshape(I, S, B))
            if dof == 6:
                out["noise_rot"] = to_np(torch.stack(tape.rot).reshape(I, S, B, 4))
    os.makedirs(os.path.join(ROOT, "tests", "golden", "fallback"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "fallback", name + ".npz")
    np.savez_compressed(path, **out)
    lt = out["ref32_mc_trans_cov_tril"][0]
    msg = f"{name}: kinds={kinds}  L_t[0] diag of the 'trans' objects: " + \
          str([np.diag(lt[b]).tolist() for b, k in enumerate(kinds) if k == "trans"])
    if dof == 6:
        lr = out["ref32_mc_rot_cov_tril"][0]
        msg += "  L_r[0] diag of the 'rot' objects: " + str([np.diag(lr[b]).tolist() for b, k in enumerate(kinds) if k == "rot"])
    print(msg, " finite logw:", bool(np.isfinite(out["ref64_mc_logw"]).all()), f"-> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(99)
    torch.set_num_threads(4)
    run_case("fallback6", 6, ["ok", "trans", "rot", "ok", "trans"])
    run_case("fallback4", 4, ["trans", "ok", "trans", "ok"])


if __name__ == "__main__":
    main()
