"""epnp_rslm_draw_f32 on a real GPU: the assertions of tests/test_rslm_draw_cpu.py (which runs them on the CPU emulation
of the kernel), plus a detection-sized launch."""
import pytest
import torch

import test_rslm_draw_cpu as _cpu

pytestmark = pytest.mark.gpu


@pytest.fixture
def dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


test_subsets_are_distinct_in_range_reproducible_and_tiling_invariant = _cpu.test_subsets_are_distinct_in_range_reproducible_and_tiling_invariant
test_first_pick_follows_the_weights = _cpu.test_first_pick_follows_the_weights
test_inclusion_frequencies_match_torch_multinomial = _cpu.test_inclusion_frequencies_match_torch_multinomial
test_too_few_positive_weights_completes_the_subset_in_index_order = _cpu.test_too_few_positive_weights_completes_the_subset_in_index_order
test_start_poses = _cpu.test_start_poses
test_centre_based_translation_guess_matches_the_class = _cpu.test_centre_based_translation_guess_matches_the_class
test_subclass_translation_guess_is_respected = _cpu.test_subclass_translation_guess_is_respected
test_bad_arguments_are_refused = _cpu.test_bad_arguments_are_refused
test_solver_with_native_draws = _cpu.test_solver_with_native_draws


def test_large_batch_shapes(dev):
    """Detection-sized launch (B = 4096, N = 512, 64 proposals x 16 points): every subset distinct and in range."""
    B, N, P, n = 4096, 512, 64, 16
    w2d = torch.rand(B, N, 2, device=dev) + 0.01
    from epropnp_b200 import native
    inds, start = native.rslm_draw(None, None, w2d, None, P, n, 6, seed=2, t_init=torch.zeros(B, 3, device=dev))
    assert inds.min() >= 0 and inds.max() < N
    s = inds.sort(dim=-1).values
    assert (s[..., 1:] != s[..., :-1]).all()
    assert torch.allclose(start[..., 3:].norm(dim=-1), torch.ones(P, B, device=dev), atol=1e-5)
    # heavier correspondences are drawn more often: rank correlation of inclusion counts with the weights
    cnt = torch.zeros(B, N, device=dev).scatter_add_(1, inds.permute(1, 0, 2).reshape(B, -1).long(), torch.ones(B, P * n, device=dev))
    wbar = w2d.mean(-1)
    corr = torch.stack([torch.corrcoef(torch.stack((cnt[b], wbar[b])))[0, 1] for b in range(0, B, 512)])
    assert (corr > 0.3).all(), corr
