#!/usr/bin/env python
"""TEST INFRASTRUCTURE ONLY.  Generates tests/golden/epilogue/mc_loss.npz by running the UNMODIFIED reference loss
(/root/reference/EPro-PnP-6DoF/lib/models/monte_carlo_pose_loss.py, imports only torch) on seeded log-weights that
include the corner cases it special-cases (a NaN object) and the infinities torch.logsumexp defines (all -inf, a +inf).

    python oracle/make_golden_mc_loss.py        # needs /root/reference; run in the build container

The detection flavour of the loss (EPro-PnP-Det/epropnp_det/models/losses/monte_carlo_pose_loss.py) and the Monte-Carlo
score (deform_pnp_head.py:524,533-536) import mmdet / mmcv, which are absent here, so they cannot be executed: their
expected values below are RESTATED from those lines (labelled `restated_*`, parity unpinned for them); the per-object loss
both flavours share is the executed one.
"""
import importlib.util
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/EPro-PnP-6DoF/lib/models/monte_carlo_pose_loss.py"
spec = importlib.util.spec_from_file_location("ref_mc_loss", REF)
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

M, B = 200, 12                                   # M deliberately not a multiple of the CTA size
g = torch.Generator().manual_seed(7)
logw = torch.randn(M, B, generator=g, dtype=torch.float64) * 3 - 5
logw[17, 5] = float("nan")                       # a NaN log-weight -> the reference zeroes that object's loss
logw[:, 9] += 80                                 # large offset: needs the max subtraction
logw_edge = logw.clone()                         # infinities: compared on lse / softmax only (the mean loss is NaN)
logw_edge[:, 3] = float("-inf")                  # every sample impossible
logw_edge[4, 7] = float("inf")
cost_target = torch.rand(B, generator=g, dtype=torch.float64) * 4
norm_in = torch.tensor(2.5, dtype=torch.float64)
coef = torch.randn((), generator=g, dtype=torch.float64)

out = {}
for name, dt in (("ref64", torch.float64), ("ref32", torch.float32)):
    lw = logw.to(dt).clone().requires_grad_(True)
    ct = cost_target.to(dt).clone().requires_grad_(True)
    mod = ref.MonteCarloPoseLoss(init_norm_factor=1.5, momentum=0.01)
    mod.train()
    loss = mod(lw, ct, norm_in.to(dt))
    (loss * coef.to(dt)).backward()
    out[f"{name}_loss"] = loss.detach().numpy()
    out[f"{name}_norm_factor_after"] = mod.norm_factor.numpy()
    out[f"{name}_grad_logw"] = lw.grad.numpy()
    out[f"{name}_grad_cost_target"] = ct.grad.numpy()
    mod.eval()
    out[f"{name}_loss_eval"] = mod(lw.detach(), ct.detach(), norm_in.to(dt)).numpy()
    out[f"{name}_lse"] = torch.logsumexp(lw.detach(), dim=0).numpy()

# restated (not executable here): Det MC score on 4DoF and 6DoF-shaped samples, deform_pnp_head.py:524,533-536
for D in (4, 7):
    samples = torch.randn(M, B, D, generator=g, dtype=torch.float64)
    pose_opt = samples.mean(0) + 0.1 * torch.randn(B, D, generator=g, dtype=torch.float64)
    samples[3, 0] = pose_opt[0]                  # zero deviation: log2(0) = -inf -> score 1
    finite = logw.clone()
    w = finite.softmax(dim=0)
    dev = (samples[..., [0, 2]] - pose_opt[:, [0, 2]]).norm(dim=-1)
    score = (((-dev.log2() + 2.5) / 4).clamp(min=0, max=1) * w).sum(dim=0)
    out[f"samples_d{D}"] = samples.numpy()
    out[f"pose_opt_d{D}"] = pose_opt.numpy()
    out[f"restated_score_te_d{D}"] = score.numpy()
out["restated_weights"] = logw.softmax(dim=0).numpy()
out["logw_edge"] = logw_edge.numpy()
out["ref64_lse_edge"] = torch.logsumexp(logw_edge, dim=0).numpy()
out["ref64_weights_edge"] = logw_edge.softmax(dim=0).numpy()
out.update(logw=logw.numpy(), cost_target=cost_target.numpy(), norm_in=norm_in.numpy(), coef=coef.numpy(),
           init_norm_factor=np.float64(1.5), momentum=np.float64(0.01))
path = os.path.join(os.path.dirname(HERE), "tests", "golden", "epilogue", "mc_loss.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items() if hasattr(v, "shape") and v.ndim})
