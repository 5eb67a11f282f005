"""Proposal distributions of the AMIS loop as standalone torch distributions (reference
epropnp/distributions.py).  The native kernels carry their own fused sampler / density; these classes
keep the public names importable and serve callers (and tests) that want the densities on their own.
No pyro dependency: they derive from torch.distributions directly."""
import math

import torch
from torch.distributions import Distribution, VonMises, constraints


class AngularCentralGaussian(Distribution):
    """ACG on S^(q-1): x = L z / |L z|, z ~ N(0, I_q)."""
    arg_constraints = {'scale_tril': constraints.lower_cholesky}
    has_rsample = True

    def __init__(self, scale_tril, validate_args=None, eps=1e-6):
        q = scale_tril.size(-1)
        assert q > 1 and scale_tril.shape[-2:] == (q, q)
        self.scale_tril = scale_tril
        self._unbroadcasted_scale_tril = scale_tril
        self.q = q
        self.area = 2 * math.pi ** (0.5 * q) / math.gamma(0.5 * q)
        self.eps = eps
        super().__init__(scale_tril.shape[:-2], (q,), validate_args=False)

    def log_prob(self, value):
        L = self.scale_tril
        shape = torch.broadcast_shapes(value.shape[:-1], L.shape[:-2])
        y = torch.linalg.solve_triangular(L.expand(shape + L.shape[-2:]),
                                          value.expand(shape + (self.q,)).unsqueeze(-1), upper=False).squeeze(-1)
        half_log_det = L.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        return y.square().sum(-1).log() * (-self.q / 2) - half_log_det - math.log(self.area)

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        z = torch.randn(shape, dtype=self.scale_tril.dtype, device=self.scale_tril.device)
        g = (self.scale_tril @ z.unsqueeze(-1)).squeeze(-1)
        norm = g.norm(dim=-1, keepdim=True)
        pole = torch.zeros_like(g)
        pole[..., 0] = 1.0
        return torch.where(norm < self.eps, pole, g / norm)


class VonMisesUniformMix(VonMises):
    """(1 - uniform_mix) von Mises + uniform_mix uniform on the circle."""

    def __init__(self, loc, concentration, uniform_mix=0.25, **kwargs):
        super(VonMisesUniformMix, self).__init__(loc, concentration, **kwargs)
        self.uniform_mix = uniform_mix

    @torch.no_grad()
    def sample(self, sample_shape=torch.Size()):
        assert len(sample_shape) == 1
        total = sample_shape[0]
        n_uniform = round(total * self.uniform_mix)
        shape_u = self._extended_shape((n_uniform,))
        uni = (torch.rand(shape_u, dtype=self.loc.dtype, device=self.loc.device) * 2 - 1) * math.pi
        vm = super(VonMisesUniformMix, self).sample((total - n_uniform,))
        return torch.cat((uni, vm), dim=0)

    def log_prob(self, value):
        vm = super(VonMisesUniformMix, self).log_prob(value) + math.log(1 - self.uniform_mix)
        flat = torch.full_like(vm, math.log(self.uniform_mix / (2 * math.pi)))
        return torch.logaddexp(vm, flat)
