"""The `cholesky_wrapper` fallback (epropnp.py:16-33: not positive definite -> diag(default) / identity), pinned to
tests/golden/fallback/*.npz from the UNMODIFIED reference layer (oracle/make_golden_fallback.py: crafted per-object
covariances, AMIS run end to end).  Here: the oracle's robust_cholesky / amis_6dof / amis_4dof against those vectors."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, err_vs
from oracle import pnp_oracle as orc

FALLBACK_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "fallback", "*.npz")))


def load_fallback(name):
    g = np.load(os.path.join(GOLDEN_DIR, "fallback", name + ".npz"))
    return {k: g[k] for k in g.files}


def fallback_noise(g, dtype, dev="cpu", object_major=False):
    dof = int(g["dof"])
    rot_key = "noise_rot" if dof == 6 else ("yaw_samples64" if dtype == torch.float64 else "yaw_samples")
    keys = ("noise_normal", "noise_chi2", rot_key)
    if not object_major:
        return tuple(torch.from_numpy(g[k]).to(dtype) for k in keys)
    B = int(g["B"])
    perm = {4: (2, 0, 1, 3), 3: (2, 0, 1)}
    out = []
    for k in keys:
        a = np.transpose(g[k], perm[g[k].ndim])
        out.append(torch.from_numpy(a.reshape((B, -1) + a.shape[3:]).copy()).to(dtype).to(dev))
    return tuple(out)


def test_cases_exist():
    assert set(FALLBACK_CASES) >= {"fallback4", "fallback6"}


@pytest.mark.parametrize("name", FALLBACK_CASES)
def test_first_proposal_uses_the_default_factor(name):
    g = load_fallback(name)
    dof = int(g["dof"])
    cov = torch.from_numpy(g["pose_cov_in"]).double()
    L = orc.robust_cholesky(cov[:, :3, :3], None if dof == 6 else [1.0, 1.0, 4.0])
    assert err_vs(L, g["ref64_mc_trans_cov_tril"][0]) < 1e-6          # (the ok objects come from float32 covariances)
    want = np.eye(3) if dof == 6 else np.diag([1.0, 1.0, 4.0])
    for b, kind in enumerate(g["kinds"]):
        if kind == "trans":
            assert np.array_equal(L[b].numpy(), want) and np.array_equal(g["ref32_mc_trans_cov_tril"][0][b], want)
        if kind == "rot":
            assert np.array_equal(g["ref32_mc_rot_cov_tril"][0][b], np.eye(4))


@pytest.mark.parametrize("name", FALLBACK_CASES)
def test_oracle_amis_through_the_fallback_fp64(name):
    g = load_fallback(name)
    d = torch.float64
    t = lambda k: torch.from_numpy(g[k]).to(d)
    cam = orc.Camera(t("cam_mats"), 0.1)
    delta = orc.adaptive_delta(t("x2d"), t("w2d"), 0.5)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    fn = orc.amis_6dof if int(g["dof"]) == 6 else orc.amis_4dof
    r = fn(t("x3d"), t("x2d"), t("w2d"), cam, delta, t("pose_opt_in"), t("pose_cov_in"), fallback_noise(g, d), M, I)
    assert err_vs(r["trans_tril"], g["ref64_mc_trans_cov_tril"]) < 1e-6
    if int(g["dof"]) == 6:
        assert err_vs(r["rot_tril"], g["ref64_mc_rot_cov_tril"]) < 1e-6
    assert err_vs(r["samples"], g["ref64_mc_samples"]) < 1e-7
    assert err_vs(r["logw"], g["ref64_mc_logw"]) < 1e-7
