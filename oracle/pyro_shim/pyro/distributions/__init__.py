"""Shim for `pyro.distributions` (see ../__init__.py).  Names used by the reference:
    MultivariateStudentT   epropnp/epropnp.py:10,224,232,306,312
    TorchDistribution      epropnp/distributions.py:11
    constraints            epropnp/distributions.py:11 (.lower_cholesky)
"""
import math

import torch
from torch.distributions import Distribution, constraints  # noqa: F401  (re-exported)

from .util import broadcast_shape


class TorchDistribution(Distribution):
    """pyro's TorchDistribution = torch Distribution + mixin conveniences; the reference only
    relies on the torch base-class behaviour (`sample` -> `rsample` under no_grad,
    `_extended_shape`, `_validate_args`)."""
    pass


# Noise taps: make_golden.py replaces these to record the base noise of every draw.
def draw_standard_normal(shape, dtype, device):
    return torch.empty(shape, dtype=dtype, device=device).normal_()


def draw_chi2(df, sample_shape):
    return torch.distributions.Chi2(df).rsample(sample_shape)


class MultivariateStudentT(TorchDistribution):
    """Multivariate Student-t with `df` degrees of freedom, location `loc` (..., n) and
    lower-triangular scale factor `scale_tril` (..., n, n).

    sample   x = loc + L (g * sqrt(df / c)),  g ~ N(0, I_n),  c ~ chi2(df)
    density  log p(x) = lgamma((df+n)/2) - lgamma(df/2) - (n/2) log(df*pi) - sum log diag L
                        - (df+n)/2 * log1p(|L^-1 (x-loc)|^2 / df)
    """
    arg_constraints = {"df": constraints.positive, "loc": constraints.real_vector,
                       "scale_tril": constraints.lower_cholesky}
    support = constraints.real_vector
    has_rsample = True

    def __init__(self, df, loc, scale_tril, validate_args=None):
        n = loc.size(-1)
        assert scale_tril.shape[-2:] == (n, n)
        if not torch.is_tensor(df):
            df = loc.new_tensor(float(df))
        batch_shape = broadcast_shape(df.shape, loc.shape[:-1], scale_tril.shape[:-2])
        self.df = df.expand(batch_shape)
        self.loc = loc.expand(batch_shape + (n,))
        self.scale_tril = scale_tril.expand(batch_shape + (n, n))
        super().__init__(batch_shape, (n,), validate_args=False)

    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        g = draw_standard_normal(shape, self.df.dtype, self.df.device)
        c = draw_chi2(self.df, sample_shape)
        y = g * torch.rsqrt(c / self.df).unsqueeze(-1)
        return self.loc + (self.scale_tril @ y.unsqueeze(-1)).squeeze(-1)

    def log_prob(self, value):
        n = self.loc.size(-1)
        diff = value - self.loc
        tril = self.scale_tril.expand(diff.shape[:-1] + (n, n))
        y = torch.linalg.solve_triangular(tril, diff.unsqueeze(-1), upper=False).squeeze(-1)
        log_norm = (self.scale_tril.diagonal(dim1=-2, dim2=-1).log().sum(-1)
                    + 0.5 * n * self.df.log() + 0.5 * n * math.log(math.pi)
                    + torch.lgamma(0.5 * self.df) - torch.lgamma(0.5 * (self.df + n)))
        return -0.5 * (self.df + n) * torch.log1p(y.pow(2).sum(-1) / self.df) - log_norm
