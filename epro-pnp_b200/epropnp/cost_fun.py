"""Huber cost objects (reference epropnp/cost_fun.py).  The solver kernels only read `.delta`
(float or (B,) tensor) and `.eps`; `compute` is the standalone utility of the reference API."""
import torch

from epropnp_b200 import native


def huber_kernel(s_sqrt, delta):
    """rho/2 of the Huber loss at residual norm `s_sqrt`   (cost_fun.py:8-12)."""
    return torch.where(s_sqrt <= delta, 0.5 * s_sqrt * s_sqrt, delta * s_sqrt - 0.5 * delta * delta)


def huber_d_kernel(s_sqrt, delta, eps: float = 1e-10):
    """sqrt(rho'): the robust rescaling factor of residual and Jacobian   (cost_fun.py:15-20)."""
    return (delta / s_sqrt.clamp(min=eps)).clamp(max=1.0).sqrt()


class HuberPnPCost(object):

    def __init__(self, delta=1.0, eps=1e-10):
        super(HuberPnPCost, self).__init__()
        self.eps = eps
        self.delta = delta

    def set_param(self, *args, **kwargs):
        pass

    def compute(self, x2d_proj, x2d, w2d, jac_cam=None, out_residual=False, out_cost=False, out_jacobian=False):
        """Weighted reprojection residual, Huber cost and robustly rescaled residual / Jacobian
        (cost_fun.py:33-89).  Shapes: x2d_proj/x2d/w2d (*, n, 2); jac_cam (*, n, 2, dof)."""
        lead, n = x2d_proj.shape[:-2], x2d_proj.size(-2)
        delta = self.delta if torch.is_tensor(self.delta) else x2d.new_tensor(self.delta)
        delta = delta[..., None]
        r = (x2d_proj - x2d) * w2d
        s = r.norm(dim=-1)
        residual = cost = jacobian = None
        if out_cost is not False:
            cost = huber_kernel(s, delta).sum(dim=-1)
            if torch.is_tensor(out_cost):
                out_cost.copy_(cost)
                cost = out_cost
        if out_residual is not False or out_jacobian is not False:
            scale = huber_d_kernel(s, delta, eps=self.eps)
            if out_residual is not False:
                residual = (r * scale[..., None]).reshape(*lead, n * 2)
                if torch.is_tensor(out_residual):
                    out_residual.view(residual.shape).copy_(residual)
                    residual = out_residual
            if out_jacobian is not False:
                assert jac_cam is not None
                dof = jac_cam.size(-1)
                jacobian = (jac_cam * (w2d * scale[..., None])[..., None]).reshape(*lead, n * 2, dof)
                if torch.is_tensor(out_jacobian):
                    out_jacobian.view(jacobian.shape).copy_(jacobian)
                    jacobian = out_jacobian
        return residual, cost, jacobian

    def _map_delta(self, fn):
        if torch.is_tensor(self.delta):
            self.delta = fn(self.delta)
        return self

    def reshape_(self, *batch_shape):
        return self._map_delta(lambda d: d.reshape(*batch_shape))

    def expand_(self, *batch_shape):
        return self._map_delta(lambda d: d.expand(*batch_shape))

    def repeat_(self, *batch_repeat):
        return self._map_delta(lambda d: d.repeat(*batch_repeat))

    def shallow_copy(self):
        return HuberPnPCost(delta=self.delta, eps=self.eps)


class AdaptiveHuberPnPCost(HuberPnPCost):

    def __init__(self, delta=None, relative_delta=0.5, eps=1e-10):
        super(HuberPnPCost, self).__init__()
        self.delta = delta
        self.relative_delta = relative_delta
        self.eps = eps

    def set_param(self, x2d, w2d):
        """delta = mean(w2d) * sqrt(sum_xy var(x2d)) * relative_delta per object   (cost_fun.py:123-126)."""
        differentiable = torch.is_grad_enabled() and (x2d.requires_grad or w2d.requires_grad)
        if x2d.is_cuda and x2d.dim() == 3 and not differentiable:
            self.delta = native.adaptive_delta(x2d, w2d, self.relative_delta).to(x2d.dtype)
        else:   # keeps the autograd graph (training) / generic batch shapes
            spread = torch.var(x2d, dim=-2).sum(dim=-1).sqrt()
            self.delta = w2d.mean(dim=(-2, -1)) * spread * self.relative_delta

    def shallow_copy(self):
        return AdaptiveHuberPnPCost(delta=self.delta, relative_delta=self.relative_delta, eps=self.eps)
