"""The native backward of pose_opt_plus on a real GPU: the assertions of tests/test_gn_plus_backward_cpu.py (which runs
them on the CPU emulation of the kernels)."""
import os

import pytest
import torch

import test_gn_plus_backward_cpu as _cpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


test_native_backward_matches_composite = _cpu.test_native_backward_matches_composite
test_layer_uses_it_when_asked = _cpu.test_layer_uses_it_when_asked
