"""Pins the oracle's random-sample LM initialiser (oracle/pnp_oracle.py: center_based_init, rslm_solve, select_start)
against tests/golden/rslm/*.npz, which oracle/make_golden_rslm.py produced by running the UNMODIFIED reference
RSLMSolver (levenberg_marquardt.py:268-353) with its random draws taped.  float64: algorithmic identity."""
import glob
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, err_vs
from oracle import pnp_oracle as orc

RSLM_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "rslm", "*.npz")))


def load_rslm(name):
    g = np.load(os.path.join(GOLDEN_DIR, "rslm", name + ".npz"))
    return {k: g[k] for k in g.files}


def rslm_setup(g, dtype):
    t = lambda k: torch.from_numpy(g[k]).to(dtype)
    lb = ub = None
    if int(g["bounds"]) == 2:
        lb, ub = t("lb"), t("ub")
    cam = orc.Camera(t("cam_mats"), float(g["z_min"]), lb, ub)
    delta = orc.adaptive_delta(t("x2d"), t("w2d"), float(g["relative_delta"]))
    return t("x3d"), t("x2d"), t("w2d"), cam, delta


def test_cases_exist():
    assert len(RSLM_CASES) >= 4


@pytest.mark.parametrize("name", RSLM_CASES)
def test_rslm_oracle_matches_reference_fp64(name):
    g = load_rslm(name)
    d = torch.float64
    x3d, x2d, w2d, cam, delta = rslm_setup(g, d)
    dof = int(g["dof"])
    # start poses: centre-based translation + the taped orientation draw
    start = torch.empty((int(g["P"]), int(g["B"]), 4 if dof == 4 else 7), dtype=d)
    start[..., :3] = orc.center_based_init(x2d, x3d, cam, dof)
    rot = torch.from_numpy(g["rot_draw"]).to(d)
    if dof == 4:
        start[..., 3] = rot * (2 * np.pi)
    else:
        start[..., 3:] = rot / rot.norm(dim=-1, keepdim=True)
    assert err_vs(start, g["ref64_start"]) < 1e-12
    prm = orc.LMParams(num_iter=int(g["rs_iter"]))
    r = orc.rslm_solve(x3d, x2d, w2d, cam, delta, torch.from_numpy(g["inds"]), start, prm, fast_mode=bool(g["fast_mode"]))
    assert err_vs(r["hyp_pose"], g["ref64_hyp_pose"]) < 1e-8
    assert err_vs(r["hyp_cost"], g["ref64_hyp_cost"]) < 1e-8
    assert (r["winner"].numpy() == g["ref64_winner"]).all()
    assert err_vs(r["pose"], g["ref64_best_pose"]) < 1e-8
    assert err_vs(r["cost"], g["ref64_min_cost"]) < 1e-8


@pytest.mark.parametrize("name", RSLM_CASES)
def test_force_init_selection_and_solve_fp64(name):
    """LMSolver.solve(force_init_solve=True): the reference draws a SECOND set of hypotheses there (force_inds /
    force_rot_draw); the cheaper of pose_init and the initialiser's winner is the start of the LM iterations."""
    g = load_rslm(name)
    d = torch.float64
    x3d, x2d, w2d, cam, delta = rslm_setup(g, d)
    dof = int(g["dof"])
    start = torch.empty((int(g["P"]), int(g["B"]), 4 if dof == 4 else 7), dtype=d)
    start[..., :3] = orc.center_based_init(x2d, x3d, cam, dof)
    rot = torch.from_numpy(g["force_rot_draw"]).to(d)
    if dof == 4:
        start[..., 3] = rot * (2 * np.pi)
    else:
        start[..., 3:] = rot / rot.norm(dim=-1, keepdim=True)
    fast = bool(g["fast_mode"])
    r = orc.rslm_solve(x3d, x2d, w2d, cam, delta, torch.from_numpy(g["force_inds"]), start,
                       orc.LMParams(num_iter=int(g["rs_iter"])), fast_mode=fast)
    assert err_vs(r["hyp_cost"], g["ref64_force_hyp_cost"]) < 1e-8
    pose_init = torch.from_numpy(g["pose_init"]).to(d)
    cost_init = orc.evaluate(x3d, x2d, w2d, pose_init, cam, delta)["cost"]
    assert err_vs(cost_init, g["ref64_force_cost_init"]) < 1e-10
    sel = orc.select_start(pose_init, cost_init, r["pose"], r["cost"])
    assert ((cost_init < r["cost"]).numpy() == g["ref64_force_use_init"]).all()
    assert err_vs(sel, g["ref64_force_pose_start"]) < 1e-8
    pose, _, cost = orc.lm_solve(x3d, x2d, w2d, cam, delta, sel, orc.LMParams(num_iter=int(g["lm_iter"])), fast_mode=fast)
    assert err_vs(pose, g["ref64_force_pose"]) < 1e-7
    assert err_vs(cost, g["ref64_force_cost"]) < 1e-7
