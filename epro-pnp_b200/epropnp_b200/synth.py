"""Seeded synthetic correspondence sets for the EPro-PnP hot path.

The distribution follows SURVEY.md section 8(d): one pin-hole camera per object
(f=800 px, principal point (320, 240)), objects 4-8 m in front of the camera,
512 (or N) 3D points in a unit cube, 2D points = projection + pixel noise, weights
that mimic the "softmax x global scale" output of the reference's demo network
(/root/reference/demo/fit_identity.ipynb, model cell) and an initial pose a few
centimetres / degrees away from the truth.

Everything is generated on the CPU with an explicit torch.Generator so the GPU box,
this container and the golden-vector script all see the same numbers.  Nothing here
touches the GPU or the oracle; it only builds inputs.
"""
import math

import torch

__all__ = ["make_problem", "make_noise", "quat_to_mat_ref"]


def quat_to_mat_ref(q):
    """(..., 4) [w, i, j, k] -> (..., 3, 3).  Plain textbook formula, used only to
    fabricate inputs (the product path has its own device implementation)."""
    w, x, y, z = q.unbind(-1)
    m = torch.stack((
        1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w),
        2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w),
        2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)), dim=-1)
    return m.reshape(q.shape[:-1] + (3, 3))


def _yaw_to_mat(yaw):
    c, s = torch.cos(yaw), torch.sin(yaw)
    z, o = torch.zeros_like(yaw), torch.ones_like(yaw)
    return torch.stack((c, z, s, z, o, z, -s, z, c), dim=-1).reshape(yaw.shape + (3, 3))


def make_problem(num_obj, num_pts, seed=0, dof=6, focal=800.0, pix_sigma_rel=0.002,
                 outlier_frac=0.0, init_trans_noise=0.05, init_rot_noise_deg=3.0,
                 grid2d=False, dtype=torch.float32):
    """Returns a dict of CPU tensors:
        x3d (B,N,3)  x2d (B,N,2)  w2d (B,N,2)  cam_mats (B,3,3)
        pose_gt (B,7|4)  pose_init (B,7|4)
    grid2d=True places x2d on a sqrt(N) x sqrt(N) pixel grid and back-projects it onto the
    object (the dense-coordinate-map case, reference EPro-PnP-6DoF/lib/test.py:157-161)."""
    g = torch.Generator().manual_seed(int(seed))
    B, N = int(num_obj), int(num_pts)
    f64 = torch.float64

    def U(*shape, lo=0.0, hi=1.0):
        return torch.rand(*shape, generator=g, dtype=f64) * (hi - lo) + lo

    def Nrm(*shape):
        return torch.randn(*shape, generator=g, dtype=f64)

    K = torch.tensor([[focal, 0.0, 320.0], [0.0, focal, 240.0], [0.0, 0.0, 1.0]], dtype=f64)
    cam = K.expand(B, 3, 3).clone()
    t = torch.cat((U(B, 2, lo=-1.0, hi=1.0), U(B, 1, lo=4.0, hi=8.0)), dim=-1)
    if dof == 6:
        q = Nrm(B, 4)
        q = q / q.norm(dim=-1, keepdim=True)
        R = quat_to_mat_ref(q)
        pose_gt = torch.cat((t, q), dim=-1)
    else:
        yaw = U(B, lo=-math.pi, hi=math.pi)
        R = _yaw_to_mat(yaw)
        pose_gt = torch.cat((t, yaw[:, None]), dim=-1)

    x3d = U(B, N, 3, lo=-0.5, hi=0.5)
    xc = x3d @ R.transpose(-1, -2) + t[:, None, :]
    xh = xc @ cam.transpose(-1, -2)
    x2d_clean = xh[..., :2] / xh[..., 2:3]
    sigma = pix_sigma_rel * focal
    x2d = x2d_clean + sigma * Nrm(B, N, 2)
    if outlier_frac > 0:
        n_out = int(round(outlier_frac * N))
        if n_out > 0:
            x2d[:, :n_out] += 40.0 * sigma * Nrm(B, n_out, 2)
    if grid2d:
        # dense coordinate map: x2d on a regular pixel grid, x3d the (noisy) back-projection
        side = int(round(math.sqrt(N)))
        assert side * side == N, "grid2d needs a square number of points"
        ys, xs = torch.meshgrid(torch.arange(side, dtype=f64), torch.arange(side, dtype=f64),
                                indexing="ij")
        grid = torch.stack((xs, ys), dim=-1).reshape(1, N, 2)
        # centre a (side x side)-pixel-step window on the projected object centre
        ctr = (t @ K.T)
        ctr = ctr[:, :2] / ctr[:, 2:3]
        step = 1.2 * focal / t[:, 2] / side          # the unit cube spans ~f/z pixels
        x2d = ctr[:, None, :] + (grid - (side - 1) / 2.0) * step[:, None, None]
        # back-project each pixel ray to depth of a random cube point, then to object frame
        depth = xc[..., 2:3]
        ray = torch.cat(((x2d - K[:2, 2]) / focal, torch.ones(B, N, 1, dtype=f64)), dim=-1)
        xc_new = ray * depth
        x3d = (xc_new - t[:, None, :]) @ R + 0.004 * Nrm(B, N, 3)

    w_raw = U(B, N, 2, lo=0.5, hi=1.5)
    w2d = w_raw / w_raw.sum(dim=1, keepdim=True) * (10.0 / sigma)

    pose_init = pose_gt.clone()
    pose_init[:, :3] += init_trans_noise * Nrm(B, 3)
    if dof == 6:
        dq = torch.cat((torch.ones(B, 1, dtype=f64),
                        0.5 * math.radians(init_rot_noise_deg) * Nrm(B, 3)), dim=-1)
        # Hamilton product q_gt * dq
        w1, x1, y1, z1 = pose_gt[:, 3:].unbind(-1)
        w2, x2, y2, z2 = dq.unbind(-1)
        qn = torch.stack((w1 * w2 - x1 * x2 - y1 * y2 - z1 * z2,
                          w1 * x2 + x1 * w2 + y1 * z2 - z1 * y2,
                          w1 * y2 - x1 * z2 + y1 * w2 + z1 * x2,
                          w1 * z2 + x1 * y2 - y1 * x2 + z1 * w2), dim=-1)
        pose_init[:, 3:] = qn / qn.norm(dim=-1, keepdim=True)
    else:
        pose_init[:, 3] += math.radians(init_rot_noise_deg) * Nrm(B)

    out = dict(x3d=x3d, x2d=x2d, w2d=w2d, cam_mats=cam, pose_gt=pose_gt, pose_init=pose_init)
    return {k: v.to(dtype).contiguous() for k, v in out.items()}


def make_noise(num_obj, mc_samples, seed=1, dof=6, dtype=torch.float32):
    """Base noise for the AMIS proposals in the kernel-native, object-major layout:
        normal3 (B, M, 3)  chi2 (B, M)  rot (B, M, 4 | 1)
    chi2 ~ chi-square(3) built as a sum of three squared normals."""
    g = torch.Generator().manual_seed(int(seed))
    B, M = int(num_obj), int(mc_samples)
    n3 = torch.randn(B, M, 3, generator=g, dtype=torch.float64)
    c = torch.randn(B, M, 3, generator=g, dtype=torch.float64).square().sum(-1)
    r = torch.randn(B, M, 4 if dof == 6 else 1, generator=g, dtype=torch.float64)
    return n3.to(dtype), c.to(dtype), r.to(dtype)
