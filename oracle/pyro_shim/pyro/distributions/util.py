"""Shim for `pyro.distributions.util` (reference use: epropnp/distributions.py:12)."""
import torch


def broadcast_shape(*shapes, **kwargs):
    return torch.broadcast_shapes(*[tuple(s) for s in shapes])
