"""The fused random-sample LM initialiser on a real GPU: the same assertions as tests/test_rslm_fused_cpu.py (which
runs them on the CPU emulation of the kernels)."""
import os

import pytest
import torch

import test_rslm_fused_cpu as _cpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


test_every_hypothesis_matches_the_unfused_path = _cpu.test_every_hypothesis_matches_the_unfused_path
test_solver_class_returns_the_cheapest_hypothesis = _cpu.test_solver_class_returns_the_cheapest_hypothesis
test_nan_hypothesis_wins_like_torch_min = _cpu.test_nan_hypothesis_wins_like_torch_min
