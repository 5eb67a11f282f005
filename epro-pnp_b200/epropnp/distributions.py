"""Stand-alone versions of the proposal families of the AMIS loop (reference epropnp/distributions.py, and the one
distribution the reference takes from pyro).

The kernels carry their own fused sampler / density for these (pnp_math.cuh: proposal_draw6 / proposal_logpdf6,
draw_yaw / proposal_logpdf4); the classes here keep the public names importable and let callers and tests use the
densities on their own (e.g. through the layer's initial_fit / gen_new_distr / estimate_params methods).  Plain
torch.distributions subclasses -- pyro is not needed.
"""
import math

import torch
from torch.distributions import Distribution, VonMises, constraints


def _forward_substitute(tril, rhs):
    """Solve tril @ y = rhs for batches that broadcast against each other."""
    lead = torch.broadcast_shapes(rhs.shape[:-1], tril.shape[:-2])
    q = tril.size(-1)
    y = torch.linalg.solve_triangular(tril.expand(lead + (q, q)), rhs.expand(lead + (q,)).unsqueeze(-1), upper=False)
    return y.squeeze(-1)


class AngularCentralGaussian(Distribution):
    """Direction of a zero-mean Gaussian: x = L z / |L z| with z ~ N(0, I_q), a density on the sphere S^(q-1):
        log p(x) = -(q/2) log(x^T (L L^T)^-1 x) - log det L - log area(S^(q-1))."""
    arg_constraints = {'scale_tril': constraints.lower_cholesky}
    has_rsample = True

    def __init__(self, scale_tril, validate_args=None, eps=1e-6):
        q = scale_tril.size(-1)
        if q < 2 or scale_tril.shape[-2:] != (q, q):
            raise AssertionError("scale_tril must be (..., q, q) with q > 1")
        self.q = q
        self.eps = eps
        self.scale_tril = scale_tril
        self._unbroadcasted_scale_tril = scale_tril
        self.area = 2 * math.pi ** (q / 2) / math.gamma(q / 2)
        super().__init__(scale_tril.shape[:-2], (q,), validate_args=False)

    def log_prob(self, value):
        whitened = _forward_substitute(self.scale_tril, value)
        log_det = self.scale_tril.diagonal(dim1=-2, dim2=-1).log().sum(-1)
        return -0.5 * self.q * whitened.square().sum(-1).log() - log_det - math.log(self.area)

    def rsample(self, sample_shape=torch.Size()):
        noise = torch.randn(self._extended_shape(sample_shape), dtype=self.scale_tril.dtype,
                            device=self.scale_tril.device)
        direction = torch.einsum('...ij,...j->...i', self.scale_tril, noise)
        length = direction.norm(dim=-1, keepdim=True)
        north_pole = torch.zeros_like(direction)
        north_pole[..., 0] = 1
        return torch.where(length < self.eps, north_pole, direction / length.clamp(min=1e-38))


class MultivariateStudentT(Distribution):
    """Multivariate Student-t: x = loc + L z sqrt(df / c), z ~ N(0, I_n), c ~ chi^2(df).  The translation proposal of
    the AMIS loop with df = 3 -- the reference imports it from pyro (epropnp.py:10, call sites :224, :306); same
    constructor arguments (df, loc, scale_tril), batch shapes broadcast.
        log p(x) = lgamma((df+n)/2) - lgamma(df/2) - (n/2) log(df pi) - log det L - ((df+n)/2) log(1 + |L^-1 (x - loc)|^2 / df)"""
    arg_constraints = {'df': constraints.positive, 'loc': constraints.real_vector, 'scale_tril': constraints.lower_cholesky}
    support = constraints.real_vector
    has_rsample = True

    def __init__(self, df, loc, scale_tril, validate_args=None):
        n = loc.size(-1)
        if scale_tril.shape[-2:] != (n, n):
            raise AssertionError("scale_tril must be (..., n, n) for a loc of (..., n)")
        self.df = torch.as_tensor(df, dtype=loc.dtype, device=loc.device)
        batch = torch.broadcast_shapes(self.df.shape, loc.shape[:-1], scale_tril.shape[:-2])
        self.loc, self.scale_tril, self.n = loc, scale_tril, n
        super().__init__(batch, (n,), validate_args=False)

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        z = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        chi2 = torch.distributions.Chi2(self.df.expand(shape[:-1])).rsample()
        y = z * torch.rsqrt(chi2 / self.df).unsqueeze(-1)
        return self.loc + torch.einsum('...ij,...j->...i', self.scale_tril, y)

    def log_prob(self, value):
        whitened = _forward_substitute(self.scale_tril, value - self.loc)
        n, df = self.n, self.df
        log_norm = (self.scale_tril.diagonal(dim1=-2, dim2=-1).log().sum(-1) + 0.5 * n * torch.log(df * math.pi)
                    + torch.lgamma(0.5 * df) - torch.lgamma(0.5 * (df + n)))
        return -0.5 * (df + n) * torch.log1p(whitened.square().sum(-1) / df) - log_norm


class VonMisesUniformMix(VonMises):
    """Yaw proposal of the 4DoF layer: with probability `uniform_mix` uniform on the circle, otherwise von Mises."""

    def __init__(self, loc, concentration, uniform_mix=0.25, **kwargs):
        super().__init__(loc, concentration, **kwargs)
        self.uniform_mix = uniform_mix

    @torch.no_grad()
    def sample(self, sample_shape=torch.Size()):
        """Stratified like the reference: the first round(n * uniform_mix) draws are uniform, the rest von Mises."""
        if len(sample_shape) != 1:
            raise AssertionError("sample_shape must be one-dimensional")
        n = sample_shape[0]
        n_flat = round(n * self.uniform_mix)
        flat_shape = self._extended_shape((n_flat,))
        flat = math.pi * (2 * torch.rand(flat_shape, dtype=self.loc.dtype, device=self.loc.device) - 1)
        peaked = VonMises.sample(self, (n - n_flat,))
        return torch.cat((flat, peaked), dim=0)

    def log_prob(self, value):
        peaked = VonMises.log_prob(self, value) + math.log1p(-self.uniform_mix)
        flat = math.log(self.uniform_mix) - math.log(2 * math.pi)
        return torch.logaddexp(peaked, torch.full_like(peaked, flat))
