"""The in-kernel Cholesky fallback (chol3_or_identity / chol3_or_default4 / acg_dispersed_chol in csrc/pnp_math.cuh)
against the UNMODIFIED reference's `cholesky_wrapper` behaviour (epropnp.py:16-33), on covariances that are not positive
definite for SOME objects of the batch (tests/golden/fallback/*.npz, oracle/make_golden_fallback.py): the AMIS kernel is
started from the crafted (pose_opt, pose_cov) with the reference's base noise."""
import numpy as np
import pytest
import torch

from conftest import err_stats, err_vs, record_parity
from epropnp_b200 import native
from test_fallback_cpu import FALLBACK_CASES, fallback_noise, load_fallback

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", FALLBACK_CASES)
def test_amis_kernel_through_the_fallback(cuda_device, name):
    g = load_fallback(name)
    dev = cuda_device
    t = lambda k: torch.from_numpy(g[k]).to(dev)
    dof, B, M, I = int(g["dof"]), int(g["B"]), int(g["mc_samples_total"]), int(g["mc_iters"])
    prob = native.Problem(t("x3d"), t("x2d"), t("w2d"), t("cam_mats"), None, None, t("delta"))
    p = native.default_params(dof, mc_samples=M, mc_iter=I)
    samples, logw, props = native.amis(prob, t("pose_opt_in"), t("pose_cov_in"), p, noise=fallback_noise(g, torch.float32, dev, True),
                                       want_proposals=True)
    pr = props.cpu().numpy()                                       # (B, I, 19): mu3, Lt6 (l00 l10 l11 l20 l21 l22), ...
    want_lt = np.array([1, 0, 1, 0, 0, 1], np.float32) if dof == 6 else np.array([1, 0, 1, 0, 0, 4], np.float32)
    for b, kind in enumerate(g["kinds"]):
        if kind == "trans":
            assert np.array_equal(pr[b, 0, 3:9], want_lt), (b, pr[b, 0, 3:9])
        if kind == "rot":
            assert np.array_equal(pr[b, 0, 9:19], np.array([1, 0, 1, 0, 0, 1, 0, 0, 0, 1], np.float32))
    ref_lt = g["ref64_mc_trans_cov_tril"]                           # (I, B, 3, 3)
    lt = np.stack([ref_lt[..., i, j] for i in range(3) for j in range(i + 1)], -1)
    assert err_vs(np.transpose(pr[:, :, 3:9], (1, 0, 2)), lt) < 2e-3
    smp = samples.transpose(0, 1).cpu().numpy()
    lw = logw.transpose(0, 1).cpu().numpy()
    assert np.isfinite(lw).all()
    floor_s = err_vs(g["ref32_mc_samples"], g["ref64_mc_samples"])
    floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
    es, ew = err_vs(smp, g["ref64_mc_samples"]), err_vs(lw, g["ref64_mc_logw"])
    assert es < max(1e-4, 5 * floor_s) and ew < max(1e-4, 5 * floor_w), (es, ew, floor_s, floor_w)
    record_parity(f"fallback/{name}", samples=err_stats(smp, g["ref64_mc_samples"]), logw=err_stats(lw, g["ref64_mc_logw"]),
                  ref32_vs_ref64_samples_rel=floor_s, ref32_vs_ref64_logw_rel=floor_w)
