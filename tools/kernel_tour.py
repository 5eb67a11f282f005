#!/usr/bin/env python
"""One launch of every kernel of the library at bench-like sizes, for per-kernel `ncu --set full` captures:

    ncu --set full --clock-control none --import-source on -k regex:<kernel> -c 1 -o gpurun_out/r2_<kernel> python tools/kernel_tour.py

Shapes: the metric's (B = 4096, N = 512, M = 512); RSLM at the 6DoF demo configuration (B = 4096, 64 proposals x 16 points)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "epro-pnp_b200"))
import torch  # noqa: E402
from epropnp_b200 import native  # noqa: E402
from epropnp_b200.synth import make_problem  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    B, N, M = 4096, 512, 512
    d = {k: v.to(dev) for k, v in make_problem(B, N, seed=7).items()}
    delta = native.adaptive_delta(d["x2d"], d["w2d"], 0.5)                                   # adaptive_delta_kernel
    prob = native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, delta)
    p = native.default_params(6, mc_samples=M, mc_iter=4)
    for rep in range(2):                                                                       # second round = warm
        out = native.lm_amis_fused(prob, d["pose_init"], p, seed=1, want_cost=True, want_plus=True)   # lm_warp_kernel, amis_kernel
        native.evaluate_cost(prob, out["pose_samples"].transpose(0, 1)[:128].contiguous(), 6, 0.1)    # cost_kernel (128 poses per object)
        native.evaluate_full(prob, d["pose_init"], 6, 0.1, 1e-10, True, True, True, True)             # evaluate_full_kernel
        native.cost_backward(prob, 6, 0.1, out["pose_samples"], torch.randn(B, M, device=dev),        # cost_backward_kernel
                             d["pose_init"].reshape(B, 1, 7), torch.randn(B, 1, device=dev))
        native.gn_plus_backward(prob, out["pose_opt"], torch.randn(B, 7, device=dev), 6, 0.1, 1e-5, 1e-10)   # gn_plus_backward_kernel
        ep = native.mc_epilogue(out["logw"], out["pose_samples"], out["pose_opt"], cost_target=torch.rand(B, device=dev),
                                want_lse=True, want_loss=True, want_weights=True, want_score=True)      # mc_epilogue_kernel
        native.mc_lse_backward(out["logw"], ep["lse"], torch.randn(B, device=dev))                     # mc_lse_backward_kernel
        P, n = 64, 16
        inds, start = native.rslm_draw(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], P, n, 6, seed=1)               # rslm_draw_kernel
        native.rslm(prob, inds, start, native.default_params(6, lm_iter=3))                            # rslm_kernel
    torch.cuda.synchronize()
    print("kernel tour finished")


if __name__ == "__main__":
    main()
