#!/usr/bin/env python
"""CPU study for DESIGN.md section 9 item 3: accuracy of the AMIS cost if the 3x4 projection K[R|t] * [X;1] ran on
tensor cores.  TF32 operands (10-bit mantissa) with fp32 accumulation are emulated in numpy; "3xTF32" is the usual
error-compensated split a = a_hi + a_lo (3 products).  Reported: error of the per-sample cost against float64 for
(a) plain fp32 FMA (what the kernel does), (b) 1xTF32, (c) 3xTF32, on the bench's synthetic distribution."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "epro-pnp_b200"))
from epropnp_b200.synth import make_problem  # noqa: E402


def tf32(x):
    """round-to-nearest-even to 10 explicit mantissa bits (drop 13)"""
    u = np.asarray(x, np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x0FFF + ((u >> 13) & 1)) & ~np.uint64(0x1FFF)
    return u.astype(np.uint32).view(np.float32)


def rot(q):
    w, x, y, z = q.T
    return np.stack([1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
                     2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
                     2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)


def cost_from_xh(xh, x2d, w2d, delta, dt):
    z = np.maximum(xh[..., 2:3], dt(0.1))
    r = (xh[..., :2] / z - x2d) * w2d
    s = np.sqrt((r * r).sum(-1))
    return np.where(s <= delta, 0.5 * s * s, delta * s - 0.5 * delta * delta).sum(-1)


def main():
    B, N, S = 16, 512, 128
    pc = {k: v.numpy() for k, v in make_problem(B, N, seed=3).items()}
    rng = np.random.default_rng(0)
    poses = np.repeat(pc["pose_gt"][:, None], S, 1).astype(np.float64)
    poses[..., :3] += 0.03 * rng.standard_normal((B, S, 3))
    q = poses[..., 3:] + 0.01 * rng.standard_normal((B, S, 4))
    poses[..., 3:] = q / np.linalg.norm(q, axis=-1, keepdims=True)
    x2d, w2d = pc["x2d"].astype(np.float64), pc["w2d"].astype(np.float64)
    delta = (w2d.mean((1, 2)) * np.sqrt(x2d.var(1, ddof=1).sum(-1)) * 0.5)[:, None, None]
    res = {"fp32 FMA": [], "1xTF32": [], "3xTF32": []}
    for b in range(B):
        K = pc["cam_mats"][b].astype(np.float64)
        R = rot(poses[b, :, 3:])
        P = np.concatenate([K @ R, (K @ poses[b, :, :3, None])], -1)            # (S, 3, 4) float64
        Xh = np.concatenate([pc["x3d"][b].astype(np.float64), np.ones((N, 1))], -1)   # (N, 4)
        ref = cost_from_xh(np.einsum("sij,nj->sni", P, Xh), x2d[b], w2d[b], delta[b], np.float64)
        P32, X32 = P.astype(np.float32), Xh.astype(np.float32)
        xh32 = np.zeros((S, N, 3), np.float32)
        for j in range(4):                                                        # fp32 accumulate, fp32 products
            xh32 += P32[:, None, :, j] * X32[None, :, j, None]
        res["fp32 FMA"].append(cost_from_xh(xh32, x2d[b].astype(np.float32), w2d[b].astype(np.float32),
                                            delta[b].astype(np.float32), np.float32) - ref)
        Ph, Xhh = tf32(P32), tf32(X32)
        Pl, Xl = tf32(P32 - Ph), tf32(X32 - Xhh)
        one = np.zeros((S, N, 3), np.float32)
        three = np.zeros((S, N, 3), np.float32)
        for j in range(4):
            hh = Ph[:, None, :, j] * Xhh[None, :, j, None]
            one += hh
            three += Pl[:, None, :, j] * Xhh[None, :, j, None] + Ph[:, None, :, j] * Xl[None, :, j, None]
        three += one
        for name, xh in (("1xTF32", one), ("3xTF32", three)):
            res[name].append(cost_from_xh(xh, x2d[b].astype(np.float32), w2d[b].astype(np.float32),
                                          delta[b].astype(np.float32), np.float32) - ref)
    scale = np.abs(ref).max()
    print(f"cost of {S} poses x {N} points per object, {B} objects; |cost| up to ~{scale:.0f}")
    for name, errs in res.items():
        e = np.abs(np.concatenate([x.ravel() for x in errs]))
        print(f"  {name:9s}  |cost - float64|: median {np.median(e):.2e}  p99 {np.percentile(e, 99):.2e}  max {e.max():.2e}")


if __name__ == "__main__":
    main()
