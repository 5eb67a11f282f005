"""The native Monte-Carlo epilogue (epnp_mc_epilogue_f32 / epnp_mc_lse_backward_f32) on a real GPU: the assertions of
tests/test_mc_epilogue_cpu.py that exercise the kernels."""
import os

import pytest
import torch

import test_mc_epilogue_cpu as _cpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


native_on = _cpu.native_on
test_native_loss_matches_the_reference_run = _cpu.test_native_loss_matches_the_reference_run
test_native_lse_and_weights_at_the_infinities = _cpu.test_native_lse_and_weights_at_the_infinities
test_native_score_te = _cpu.test_native_score_te
test_detection_flavour_reductions = _cpu.test_detection_flavour_reductions
test_epilogue_argument_checks = _cpu.test_epilogue_argument_checks
