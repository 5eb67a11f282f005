#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY.  One CPU worker of bench.py's reference arm / cpu_baseline leg.

    python oracle/ref_worker.py --objects 16 --threads 4 --seconds 10 [--slices 3 --points 512 --samples 512 --config fused]

Runs the hot path on a bounded sample (`--objects` synthetic objects per pass, passes repeated for `--seconds` per
slice, `--slices` slices back to back) and prints one JSON line {slices: [[objects, seconds], ...], kind, threads}.  kind = "reference+shim": the UNMODIFIED reference layer
(oracle/_ref, staged by oracle/stage_ref.py) through its stock code path -- EProPnP6DoF.monte_carlo_forward with
LMSolver / AdaptiveHuberPnPCost / PerspectiveCamera, pyro's MultivariateStudentT served by oracle/pyro_shim;
kind = "port": oracle/pnp_oracle.py (the pinned restatement) when the reference has not been staged.
bench.py starts one worker per group of host cores (the objects are independent) and adds the rates up.
"""
import argparse
import json
import os
import sys
import time
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--objects", type=int, default=16)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--seconds", type=float, default=10.0)
    ap.add_argument("--points", type=int, default=512)
    ap.add_argument("--samples", type=int, default=512)
    ap.add_argument("--mc-iter", type=int, default=4)
    ap.add_argument("--lm-iter", type=int, default=10)
    ap.add_argument("--config", default="fused", choices=["fused", "lm_only"])
    ap.add_argument("--seed", type=int, default=5)
    ap.add_argument("--slices", type=int, default=1)
    ap.add_argument("--fast-mode", action="store_true")
    ap.add_argument("--z-min", type=float, default=0.1)
    ap.add_argument("--rel-delta", type=float, default=0.5)
    ap.add_argument("--grid2d", action="store_true")
    ap.add_argument("--force-port", action="store_true")
    a = ap.parse_args()
    os.environ.setdefault("OMP_NUM_THREADS", str(a.threads))
    warnings.filterwarnings("ignore")
    import torch
    torch.set_num_threads(a.threads)
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "epro-pnp_b200"))
    from epropnp_b200.synth import make_problem
    sys.path.remove(os.path.join(ROOT, "epro-pnp_b200"))      # `epropnp` must resolve to the staged reference below
    pc = make_problem(a.objects, a.points, seed=a.seed, grid2d=a.grid2d)
    from oracle.stage_ref import staged_path
    ref = None if a.force_port else staged_path()
    if ref:
        sys.path.insert(0, os.path.join(HERE, "pyro_shim"))
        sys.path.insert(0, ref)
        for m in [m for m in sys.modules if m == "epropnp" or m.startswith("epropnp.")]:
            del sys.modules[m]
        from epropnp.camera import PerspectiveCamera
        from epropnp.cost_fun import AdaptiveHuberPnPCost
        from epropnp.epropnp import EProPnP6DoF
        from epropnp.levenberg_marquardt import LMSolver
        import epropnp.epropnp as _m
        assert os.path.abspath(_m.__file__).startswith(os.path.abspath(ref)), _m.__file__
        camera = PerspectiveCamera(cam_mats=pc["cam_mats"], z_min=a.z_min)
        cost_fun = AdaptiveHuberPnPCost(relative_delta=a.rel_delta)
        solver = LMSolver(dof=6, num_iter=a.lm_iter)
        layer = EProPnP6DoF(mc_samples=a.samples, num_iter=a.mc_iter, solver=solver)

        def run():
            with torch.no_grad():
                cost_fun.set_param(pc["x2d"], pc["w2d"])
                if a.config == "lm_only":
                    return solver(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun, pose_init=pc["pose_init"], with_pose_cov=True,
                                  fast_mode=a.fast_mode)
                return layer.monte_carlo_forward(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun,
                                                 pose_init=pc["pose_init"], force_init_solve=False, fast_mode=a.fast_mode)
        kind = "reference+shim"
    else:
        from oracle import pnp_oracle as orc
        from epropnp_b200.synth import make_noise
        n3, c2, n4 = make_noise(a.objects, a.samples, seed=6)
        I, S, B = a.mc_iter, a.samples // a.mc_iter, a.objects
        noise = (n3.reshape(B, I, S, 3).permute(1, 2, 0, 3).contiguous(), c2.reshape(B, I, S).permute(1, 2, 0).contiguous(),
                 n4.reshape(B, I, S, 4).permute(1, 2, 0, 3).contiguous())
        cam = orc.Camera(pc["cam_mats"], a.z_min)

        def run():
            with torch.no_grad():
                delta = orc.adaptive_delta(pc["x2d"], pc["w2d"], a.rel_delta)
                if a.config == "lm_only":
                    return orc.lm_solve(pc["x3d"], pc["x2d"], pc["w2d"], cam, delta, pc["pose_init"], orc.LMParams(num_iter=a.lm_iter),
                                        fast_mode=a.fast_mode)
                return orc.monte_carlo_forward_6dof(pc["x3d"], pc["x2d"], pc["w2d"], cam, delta, pc["pose_init"], noise,
                                                    a.samples, I, orc.LMParams(num_iter=a.lm_iter), fast_mode=a.fast_mode)
        kind = "port"
    run()                                   # untimed: allocator, thread pool
    slices = []
    for _ in range(a.slices):
        t0 = time.perf_counter()
        passes = 0
        while True:
            run()
            passes += 1
            el = time.perf_counter() - t0
            if el >= a.seconds or passes >= 100000:
                break
        slices.append([passes * a.objects, el])
    print(json.dumps(dict(slices=slices, kind=kind, threads=a.threads)), flush=True)


if __name__ == "__main__":
    main()
