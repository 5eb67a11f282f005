"""PerspectiveCamera: parameter holder + standalone projection (reference epropnp/camera.py).

The solver kernels read `cam_mats`, `z_min`, `lb`, `ub` straight from this object; `project` itself
is kept for callers that want projections / camera Jacobians outside the solve (it is a handful of
torch ops and is not what LMSolver / EProPnP* execute).
"""
import torch

from .builder import CAMERA
from .common import _pose_rot, skew


def project_a(x3d, pose, cam_mats, z_min: float):
    """Rotate, translate, apply K, clamp the depth: -> x2d_proj (*, n, 2), x3d_rot (*, n, 3), z (*, n, 1)
    (reference camera.py:10-18; the form that also yields what the Jacobian needs)."""
    x3d_rot = x3d @ _pose_rot(pose).transpose(-1, -2)
    xh = (x3d_rot + pose[..., None, :3]) @ cam_mats.transpose(-1, -2)
    z = xh[..., 2:3].clamp(min=z_min)
    return xh[..., :2] / z, x3d_rot, z


def project_b(x3d, pose, cam_mats, z_min: float):
    """The same projection with K R and K t multiplied out first (cost-only evaluations): -> x2d_proj, z
    (reference camera.py:21-30)."""
    xh = x3d @ (cam_mats @ _pose_rot(pose)).transpose(-1, -2) + (cam_mats @ pose[..., :3, None]).transpose(-1, -2)
    z = xh[..., 2:3].clamp(min=z_min)
    return xh[..., :2] / z, z


@CAMERA.register_module()
class PerspectiveCamera(object):

    def __init__(self, cam_mats=None, z_min=0.1, img_shape=None, allowed_border=200, lb=None, ub=None):
        """cam_mats (*, 3, 3); img_shape (*, 2) as [h, w] or None; lb / ub: None | float | (*, 2) in [x, y]."""
        super(PerspectiveCamera, self).__init__()
        self.z_min = z_min
        self.allowed_border = allowed_border
        self.set_param(cam_mats, img_shape, lb, ub)

    def set_param(self, cam_mats, img_shape=None, lb=None, ub=None):
        self.cam_mats = cam_mats
        if img_shape is None:
            self.lb, self.ub = lb, ub
        else:   # bounds derived from the image size plus a border (reference camera.py:55-59)
            self.lb = -0.5 - self.allowed_border
            self.ub = img_shape[..., [1, 0]] + (-0.5 + self.allowed_border)

    # ------------------------------------------------------------------ projection utilities
    def _clamp(self, u):
        lb, ub = self.lb, self.ub
        if lb is None or ub is None:
            return u, None, None
        lb_t = lb.unsqueeze(-2) if torch.is_tensor(lb) else u.new_tensor(lb)
        ub_t = ub.unsqueeze(-2) if torch.is_tensor(ub) else u.new_tensor(ub)
        return torch.minimum(torch.maximum(u, lb_t), ub_t), lb_t, ub_t

    def project(self, x3d, pose, out_jac=False, clip_jac=True):
        """x3d (*, n, 3), pose (*, 4|7) -> x2d_proj (*, n, 2), jac (*, n, 2, 4|6) | None."""
        x_rot = x3d @ _pose_rot(pose).transpose(-1, -2)
        xh = (x_rot + pose[..., None, :3]) @ self.cam_mats.transpose(-1, -2)
        z = xh[..., 2:3].clamp(min=self.z_min)
        u, lb_t, ub_t = self._clamp(xh[..., :2] / z)
        if out_jac is False:
            return u, None
        jac = self.project_jacobian(x_rot, z, u, out_jac=None, dof=4 if pose.size(-1) == 4 else 6)
        if clip_jac:
            dead = (z == self.z_min).expand(u.shape)
            if lb_t is not None:
                dead = dead | (u == lb_t) | (u == ub_t)
            jac = jac.masked_fill(dead[..., None], 0)
        if torch.is_tensor(out_jac):
            out_jac.copy_(jac)
            jac = out_jac
        return u, jac

    def project_jacobian(self, x3d_rot, zcam, x2d_proj, out_jac, dof):
        if dof not in (4, 6):
            raise ValueError('dof must be 4 or 6')
        K = self.cam_mats[..., None, :, :]
        j3 = torch.cat((K[..., :2, :2] / zcam.unsqueeze(-1),
                        (K[..., :2, 2:3] - x2d_proj.unsqueeze(-1)) / zcam.unsqueeze(-1)), dim=-1)
        if dof == 6:
            j_rot = j3 @ skew(2 * x3d_rot)
        else:
            j_rot = j3[..., 0:1] * x3d_rot[..., None, 2:3] - j3[..., 2:3] * x3d_rot[..., None, 0:1]
        jac = torch.cat((j3, j_rot), dim=-1)
        if out_jac is not None:
            out_jac.copy_(jac)
            jac = out_jac
        return jac

    @staticmethod
    def get_quaternion_transfrom_mat(quaternions):
        """(*, 4) -> (*, 4, 3): maps a rotation increment in the tangent space to R^4 (camera.py:145-165)."""
        w, x, y, z = quaternions.unbind(-1)
        t = torch.stack((x, y, z, -w, -z, y, z, -w, -x, -y, x, -w), dim=-1)
        return t.reshape(quaternions.shape[:-1] + (4, 3))

    # ------------------------------------------------------------------ batch helpers (camera.py:167-197)
    def _map_bounds(self, fn):
        if torch.is_tensor(self.lb):
            self.lb = fn(self.lb)
        if torch.is_tensor(self.ub):
            self.ub = fn(self.ub)
        return self

    def reshape_(self, *batch_shape):
        self.cam_mats = self.cam_mats.reshape(*batch_shape, 3, 3)
        return self._map_bounds(lambda b: b.reshape(*batch_shape, 2))

    def expand_(self, *batch_shape):
        self.cam_mats = self.cam_mats.expand(*batch_shape, -1, -1)
        return self._map_bounds(lambda b: b.expand(*batch_shape, -1))

    def repeat_(self, *batch_repeat):
        self.cam_mats = self.cam_mats.repeat(*batch_repeat, 1, 1)
        return self._map_bounds(lambda b: b.repeat(*batch_repeat, 1))

    def shallow_copy(self):
        return PerspectiveCamera(cam_mats=self.cam_mats, z_min=self.z_min,
                                 allowed_border=self.allowed_border, lb=self.lb, ub=self.ub)
