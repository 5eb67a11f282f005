#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Golden vectors for the random-sample LM initialiser, produced by running the
UNMODIFIED reference `RSLMSolver.solve` / `LMSolver.solve(force_init_solve=True)`
(/root/reference/epropnp/levenberg_marquardt.py:80-130, 268-353, torch CPU) on seeded inputs.

    python oracle/make_golden_rslm.py            # needs /root/reference; run in the build container

The initialiser draws random numbers three ways (torch.multinomial :306, torch.randn :320 / torch.rand :317).
They are taped in call order while the float32 reference runs, stored next to the outputs, and played back for
the float64 re-run, exactly as make_golden.py does for the AMIS noise.  Stored per case (ref32_* / ref64_*):

    inds (P, B, n)            the sampled correspondence indices within each object (the same for both runs)
    start (P, B, D)           the starting hypotheses (centre-based translation + random orientation)
    hyp_pose (P, B, D)        every hypothesis after its LM / GN iterations on its n-point mini-problem
    hyp_cost (P, B)           its cost on the FULL correspondence set
    best_pose (B, D), min_cost (B), winner (B)     what RSLMSolver.solve returns (+ the arg-min)
    force_pose_start (B, D), force_use_init (B)    the start LMSolver.solve picks with force_init_solve=True
    force_pose (B, D), force_cost (B)              and where its own LM iterations end from there

hyp_pose / hyp_cost are what the reference computes inside `solve`; they are captured by wrapping the module-level
`evaluate_pnp` name the reference calls at :347 (the arguments pass through untouched).
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
sys.path.append(os.path.join(ROOT, "epro-pnp_b200"))
warnings.filterwarnings("ignore")

import epropnp.levenberg_marquardt as ref_lm             # noqa: E402  (reference)
from epropnp.camera import PerspectiveCamera             # noqa: E402
from epropnp.cost_fun import AdaptiveHuberPnPCost        # noqa: E402
from epropnp_b200.synth import make_problem              # noqa: E402

assert "/root/reference" in os.path.abspath(ref_lm.__file__)


class DrawTape:
    """Record / play back torch.multinomial, torch.randn and torch.rand in call order."""

    def __init__(self):
        self.rec = dict(multinomial=[], randn=[], rand=[])
        self.play = None
        self._orig = {}

    def __enter__(self):
        tape = self
        for name in ("multinomial", "randn", "rand"):
            orig = getattr(torch, name)
            self._orig[name] = orig

            def wrapped(*a, _name=name, _orig=orig, **kw):
                if tape.play is not None:
                    x = tape.play[_name].pop(0)
                    return x.clone() if _name == "multinomial" else x.to(kw.get("dtype", x.dtype)).clone()
                x = _orig(*a, **kw)
                tape.rec[_name].append(x.clone())
                return x
            setattr(torch, name, wrapped)
        return self

    def __exit__(self, *exc):
        for name, orig in self._orig.items():
            setattr(torch, name, orig)

    def start_playback(self):
        self.play = {k: [x.clone() for x in v] for k, v in self.rec.items()}


def to_np(x):
    return x.detach().cpu().numpy()


def run_case(name, B, N, dof, seed, P, n, rs_iter=3, lm_iter=5, fast_mode=False, outlier_frac=0.0, bounds=False,
             init_noise=(0.3, 25.0), exact_init_even=False):
    p = make_problem(B, N, seed=seed, dof=dof, outlier_frac=outlier_frac, init_trans_noise=init_noise[0],
                     init_rot_noise_deg=init_noise[1])
    if exact_init_even:          # even objects start AT the ground truth: pose_init beats the initialiser there (use_init)
        p["pose_init"][0::2] = p["pose_gt"][0::2]
    out = dict(B=B, N=N, dof=dof, P=P, n=n, rs_iter=rs_iter, lm_iter=lm_iter, fast_mode=int(fast_mode),
               relative_delta=0.5, z_min=0.1, bounds=2 if bounds else 0)
    for k in ("x3d", "x2d", "w2d", "cam_mats", "pose_init", "pose_gt"):
        out[k] = to_np(p[k])
    captured = {}
    orig_eval = ref_lm.evaluate_pnp

    def tap_eval(x3d, x2d, w2d, pose, *a, **kw):
        r = orig_eval(x3d, x2d, w2d, pose, *a, **kw)
        if pose.dim() == 3:                       # the (P, B, D) scoring call of RSLMSolver.solve (:347)
            captured["hyp_pose"], captured["hyp_cost"] = pose.clone(), r[1].clone()
        return r

    with DrawTape() as tape:
        ref_lm.evaluate_pnp = tap_eval
        try:
            for tag, dtype in (("ref32", torch.float32), ("ref64", torch.float64)):
                if tag == "ref64":
                    tape.start_playback()
                x3d, x2d, w2d = (p[k].to(dtype) for k in ("x3d", "x2d", "w2d"))
                lb = ub = None
                if bounds:
                    lb = torch.stack((p["x2d"][..., 0].min(1).values + 8, p["x2d"][..., 1].min(1).values + 8), -1).to(dtype)
                    ub = torch.stack((p["x2d"][..., 0].max(1).values - 8, p["x2d"][..., 1].max(1).values - 8), -1).to(dtype)
                    if tag == "ref32":
                        out["lb"], out["ub"] = to_np(lb), to_np(ub)
                camera = PerspectiveCamera(cam_mats=p["cam_mats"].to(dtype), z_min=0.1, lb=lb, ub=ub)
                cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
                cost_fun.set_param(x2d, w2d)
                if tag == "ref32":
                    out["delta"] = to_np(cost_fun.delta.reshape(B))
                rs = ref_lm.RSLMSolver(dof=dof, num_points=n, num_proposals=P, num_iter=rs_iter)
                # 1) the initialiser on its own
                n_mult, n_randn, n_rand = (len(tape.rec[k]) for k in ("multinomial", "randn", "rand"))
                best_pose, _, min_cost = rs.solve(x3d, x2d, w2d, camera, cost_fun, fast_mode=fast_mode)
                out[f"{tag}_hyp_pose"], out[f"{tag}_hyp_cost"] = to_np(captured["hyp_pose"]), to_np(captured["hyp_cost"])
                out[f"{tag}_best_pose"], out[f"{tag}_min_cost"] = to_np(best_pose), to_np(min_cost)
                out[f"{tag}_winner"] = to_np(captured["hyp_cost"].min(dim=0).indices)
                # the start poses: center_based_init + the taped orientation draw, as :314-324 builds them
                start = x2d.new_empty((P, B, 4 if dof == 4 else 7))
                start[..., :3] = rs.center_based_init(x2d, x3d, camera)
                if tag == "ref32":
                    inds = tape.rec["multinomial"][n_mult].reshape(P, B, n)
                    out["inds"] = to_np(inds).astype(np.int32)
                    rot = (tape.rec["rand"][n_rand] if dof == 4 else tape.rec["randn"][n_randn])
                    out["rot_draw"] = to_np(rot)
                rot_d = torch.from_numpy(out["rot_draw"]).to(dtype)
                if dof == 4:
                    start[..., 3] = rot_d * (2 * np.pi)
                else:
                    qn = rot_d.norm(dim=-1, keepdim=True)
                    start[..., 3:] = rot_d / qn
                    assert (qn >= rs.eps).all()
                out[f"{tag}_start"] = to_np(start)
                # 2) through LMSolver.solve(force_init_solve=True): use_init selection (:115-130) + its own iterations
                solver = ref_lm.LMSolver(dof=dof, num_iter=lm_iter, init_solver=rs)
                pose_init = p["pose_init"].to(dtype)
                pose_opt, _, cost = solver.solve(x3d, x2d, w2d, camera, cost_fun, pose_init=pose_init,
                                                 with_cost=True, force_init_solve=True, fast_mode=fast_mode)
                cost_init = orig_eval(x3d, x2d, w2d, pose_init, camera, cost_fun, out_cost=True)[1]
                cost_rs = captured["hyp_cost"].min(dim=0).values
                use_init = cost_init < cost_rs
                winner2 = captured["hyp_cost"].min(dim=0).indices
                start2 = captured["hyp_pose"][winner2, torch.arange(B)]
                start2[use_init] = pose_init[use_init]
                out[f"{tag}_force_hyp_cost"] = to_np(captured["hyp_cost"])
                out[f"{tag}_force_hyp_pose"] = to_np(captured["hyp_pose"])
                out[f"{tag}_force_use_init"] = to_np(use_init)
                out[f"{tag}_force_pose_start"] = to_np(start2)
                out[f"{tag}_force_cost_init"] = to_np(cost_init)
                out[f"{tag}_force_pose"], out[f"{tag}_force_cost"] = to_np(pose_opt), to_np(cost)
                if tag == "ref32":
                    out["force_inds"] = to_np(tape.rec["multinomial"][n_mult + 1].reshape(P, B, n)).astype(np.int32)
                    out["force_rot_draw"] = to_np(tape.rec["rand"][n_rand + 1] if dof == 4 else tape.rec["randn"][n_randn + 1])
        finally:
            ref_lm.evaluate_pnp = orig_eval
    os.makedirs(os.path.join(ROOT, "tests", "golden", "rslm"), exist_ok=True)
    path = os.path.join(ROOT, "tests", "golden", "rslm", name + ".npz")
    np.savez_compressed(path, **out)
    a, b = out["ref32_hyp_cost"], out["ref64_hyp_cost"]
    relc = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
    same = (out["ref32_winner"] == out["ref64_winner"]).mean()
    print(f"{name}: B={B} N={N} dof={dof} P={P} n={n}  hyp_cost 32-vs-64 rel p50 {np.median(relc):.1e} p90 {np.quantile(relc, .9):.1e} "
          f"max {relc.max():.1e}  same winner {same:.0%}  use_init {out['ref64_force_use_init'].mean():.0%}"
          f"  -> {os.path.getsize(path) / 1024:.0f} KiB")


def main():
    torch.manual_seed(4321)
    torch.set_num_threads(4)
    run_case("rslm6_basic", B=6, N=128, dof=6, seed=41, P=64, n=16)
    run_case("rslm6_small", B=5, N=64, dof=6, seed=42, P=32, n=8, outlier_frac=0.1, bounds=True, exact_init_even=True)
    run_case("rslm6_fast", B=4, N=96, dof=6, seed=43, P=32, n=16, fast_mode=True)
    run_case("rslm4_basic", B=6, N=64, dof=4, seed=44, P=64, n=16, exact_init_even=True)


if __name__ == "__main__":
    main()
