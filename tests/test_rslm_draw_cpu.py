"""The random draws of the random-sample initialiser as one kernel (epnp_rslm_draw_f32) -- weighted subsets without
replacement and uniformly random start orientations (reference: torch.multinomial / randn / rand,
levenberg_marquardt.py:306-324).  The draws come from a different generator than torch's, so parity is statistical:
the same distribution (checked against exact probabilities and against torch.multinomial's own frequencies), plus the
hard properties (distinct indices, zero weights never drawn, reproducible, independent of the batch tiling).
Runs the real kernel source under the CPU SIMT emulator; tests/test_rslm_draw_gpu.py repeats it on the GPU."""
import math

import numpy as np
import pytest
import torch

import simt_native
from epropnp.camera import PerspectiveCamera
from epropnp.common import evaluate_pnp
from epropnp.cost_fun import AdaptiveHuberPnPCost
from epropnp.levenberg_marquardt import LMSolver, RSLMSolver
from epropnp_b200 import native
from epropnp_b200.synth import make_problem
from test_oracle_rslm_cpu import RSLM_CASES, load_rslm


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


def _weights(B, N, seed, zero_every=0):
    g = torch.Generator().manual_seed(seed)
    w = torch.rand(B, N, 2, generator=g) ** 2 + 0.02
    if zero_every:
        w[:, ::zero_every] = 0.0
    return w


def test_subsets_are_distinct_in_range_reproducible_and_tiling_invariant(dev):
    B, N, P, n = 3, 37, 200, 6
    w2d = _weights(B, N, 1, zero_every=5).to(dev)
    t0 = torch.arange(3 * B, dtype=torch.float32).reshape(B, 3).to(dev)
    inds, start = native.rslm_draw(None, None, w2d, None, P, n, 6, seed=11, t_init=t0)
    assert inds.shape == (P, B, n) and inds.dtype == torch.int32 and start.shape == (P, B, 7)
    i = inds.cpu().numpy()
    assert i.min() >= 0 and i.max() < N
    assert all(len(set(row)) == n for row in i.reshape(-1, n))               # without replacement
    assert (i % 5 != 0).all()                                                 # zero-weight correspondences are never drawn
    inds2, start2 = native.rslm_draw(None, None, w2d, None, P, n, 6, seed=11, t_init=t0)
    assert torch.equal(inds, inds2) and torch.equal(start, start2)            # counter-based: same seed, same draws
    inds3, _ = native.rslm_draw(None, None, w2d, None, P, n, 6, seed=12, t_init=t0)
    assert not torch.equal(inds, inds3)
    # object 1 drawn alone with its global index == object 1 drawn inside the batch (shards draw what the batch would)
    alone_i, alone_s = native.rslm_draw(None, None, w2d[1:2], None, P, n, 6, seed=11, obj_offset=1, t_init=t0[1:2])
    assert torch.equal(alone_i[:, 0], inds[:, 1]) and torch.equal(alone_s[:, 0], start[:, 1])
    # fewer proposals: a prefix of the same streams
    few_i, few_s = native.rslm_draw(None, None, w2d, None, 7, n, 6, seed=11, t_init=t0)
    assert torch.equal(few_i, inds[:7]) and torch.equal(few_s, start[:7])


def test_first_pick_follows_the_weights(dev):
    """n = 1: P(i) = w_i / sum(w) exactly (the minimum of independent exponentials with rates w_i)."""
    N, P = 9, 8192
    w2d = _weights(1, N, 2).to(dev)
    inds, _ = native.rslm_draw(None, None, w2d, None, P, 1, 6, seed=5, t_init=torch.zeros(1, 3, device=dev))
    wbar = w2d[0].mean(-1).double().cpu().numpy()
    expect = wbar / wbar.sum()
    freq = np.bincount(inds.cpu().numpy().reshape(-1), minlength=N) / P
    sigma = np.sqrt(expect * (1 - expect) / P)
    assert (np.abs(freq - expect) < 4.5 * sigma).all(), (freq, expect)


def test_inclusion_frequencies_match_torch_multinomial(dev):
    """n of N without replacement: per-index inclusion frequency against torch.multinomial's on the same weights."""
    N, P, n = 12, 6000, 4
    w2d = _weights(1, N, 3).to(dev)
    inds, _ = native.rslm_draw(None, None, w2d, None, P, n, 6, seed=9, t_init=torch.zeros(1, 3, device=dev))
    ours = np.bincount(inds.cpu().numpy().reshape(-1), minlength=N) / P
    g = torch.Generator().manual_seed(0)
    ref_n = 60000
    ref = torch.multinomial(w2d[0].mean(-1).cpu().double().expand(ref_n, N), n, generator=g)
    theirs = np.bincount(ref.numpy().reshape(-1), minlength=N) / ref_n
    sigma = np.sqrt(theirs * (1 - theirs) * (1 / P + 1 / ref_n))
    assert (np.abs(ours - theirs) < 4.5 * sigma).all(), (ours, theirs)
    assert abs(ours.sum() - n) < 1e-9


def test_too_few_positive_weights_completes_the_subset_in_index_order(dev):
    """torch.multinomial raises here; the kernel cannot, and returns the positive ones plus the first unused indices."""
    N, n = 10, 5
    w2d = torch.zeros(1, N, 2)
    w2d[0, [3, 7]] = 1.0
    inds, _ = native.rslm_draw(None, None, w2d.to(dev), None, 16, n, 4, seed=1, t_init=torch.zeros(1, 3, device=dev))
    for row in inds.cpu().numpy().reshape(-1, n):
        assert set(row) == {3, 7, 0, 1, 2}


@pytest.mark.parametrize("dof", [6, 4])
def test_start_poses(dev, dof):
    B, P = 2, 4096
    t0 = torch.tensor([[0.1, -0.2, 3.0], [1.0, 2.0, 8.0]]).to(dev)
    _, start = native.rslm_draw(None, None, _weights(B, 8, 4).to(dev), None, P, 2, dof, seed=3, t_init=t0)
    s = start.cpu().double()
    assert torch.equal(start[..., :3], t0.expand(P, B, 3))
    if dof == 4:
        yaw = s[..., 3]
        assert (yaw >= 0).all() and (yaw < 2 * math.pi + 1e-6).all()
        assert abs(yaw.mean().item() - math.pi) < 4.5 * (2 * math.pi / math.sqrt(12)) / math.sqrt(P * B)
        assert abs(torch.cos(yaw).mean().item()) < 4.5 / math.sqrt(2 * P * B)
    else:
        q = s[..., 3:]
        assert torch.allclose(q.norm(dim=-1), torch.ones(P, B, dtype=torch.float64), atol=1e-6)
        # uniform on the 3-sphere: E[q] = 0, E[q q^T] = I / 4
        assert (q.reshape(-1, 4).mean(0).abs() < 4.5 * 0.5 / math.sqrt(P * B)).all()
        second = (q.reshape(-1, 4).T @ q.reshape(-1, 4)) / (P * B)
        assert (second - torch.eye(4, dtype=torch.float64) / 4).abs().max() < 0.02


@pytest.mark.parametrize("dof,N", [(6, 64), (4, 51), (6, 700)])
def test_centre_based_translation_guess_matches_the_class(dev, dof, N):
    """The in-kernel guess (two block reductions) against RSLMSolver.center_based_init, the torch restatement of
    levenberg_marquardt.py:283-298 (itself pinned to the reference by tests/test_host_logic_cpu.py)."""
    B = 5
    pc = {k: v.to(dev) for k, v in make_problem(B, N, seed=70 + N, dof=dof).items()}
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"])
    want = RSLMSolver(dof=dof).center_based_init(pc["x2d"].double().cpu(), pc["x3d"].double().cpu(),
                                                 PerspectiveCamera(cam_mats=pc["cam_mats"].double().cpu()))
    inds, start, t = native.rslm_draw(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], 3, 4, dof, seed=1, want_t=True)
    assert t.shape == (B, 3)
    assert torch.allclose(t.double().cpu(), want, rtol=2e-5, atol=1e-6), (t, want)
    assert torch.equal(start[..., :3], t.expand(3, B, 3))
    # a given t_init overrides it, draw for draw the same subsets and orientations
    inds2, start2 = native.rslm_draw(None, None, pc["w2d"], None, 3, 4, dof, seed=1, t_init=t + 1.0)
    assert torch.equal(inds, inds2) and torch.equal(start2[..., 3:], start[..., 3:]) and torch.equal(start2[..., :3], (t + 1.0).expand(3, B, 3))


@pytest.mark.parametrize("name", RSLM_CASES)
def test_translation_guess_against_the_unmodified_reference(dev, name):
    """tests/golden/rslm/*.npz hold the start poses the UNMODIFIED reference built (center_based_init,
    levenberg_marquardt.py:283-298, fp64 and fp32 runs).  The class's torch restatement must reproduce the fp64 values
    exactly, the kernel's fp32 guess to fp32 rounding -- no further from them than the reference's own fp32 run."""
    g = load_rslm(name)
    dof = int(g["dof"])
    want = torch.from_numpy(g["ref64_start"][0, :, :3])
    t64 = lambda k: torch.from_numpy(g[k]).double()
    got64 = RSLMSolver(dof=dof).center_based_init(t64("x2d"), t64("x3d"), PerspectiveCamera(cam_mats=t64("cam_mats")))
    assert (got64 - want).abs().max() < 1e-12 * want.abs().max()
    t32 = lambda k: torch.from_numpy(g[k]).float().to(dev)
    _, _, t = native.rslm_draw(t32("x3d"), t32("x2d"), t32("w2d"), t32("cam_mats"), 2, 2, dof, seed=1, want_t=True)
    err = (t.double().cpu() - want).abs().max() / want.abs().max()
    floor = (torch.from_numpy(g["ref32_start"][0, :, :3]).double() - want).abs().max() / want.abs().max()
    assert err < max(2e-6, 4 * floor), (float(err), float(floor))


def test_subclass_translation_guess_is_respected(dev):
    class Fixed(RSLMSolver):
        def center_based_init(self, x2d, x3d, camera, eps=1e-6):
            return x2d.new_tensor([0.0, 0.0, 5.0]).expand(x2d.shape[0], 3)

    seen = {}
    real = native.rslm_draw

    def spy(*a, **k):
        seen["t_init"] = k.get("t_init")
        return real(*a, **k)

    pc = {k: v.to(dev) for k, v in make_problem(2, 32, seed=8).items()}
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"])
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(pc["x2d"], pc["w2d"])
    native.rslm_draw = spy
    try:
        Fixed(dof=6, num_points=6, num_proposals=4, num_iter=2).solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
        assert seen["t_init"] is not None and torch.equal(seen["t_init"].cpu(), torch.tensor([[0.0, 0.0, 5.0]] * 2))
        RSLMSolver(dof=6, num_points=6, num_proposals=4, num_iter=2).solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
        assert seen["t_init"] is None
    finally:
        native.rslm_draw = real


def test_bad_arguments_are_refused(dev):
    w2d = _weights(2, 8, 5).to(dev)
    with pytest.raises(native.NativeError):
        native.rslm_draw(None, None, w2d, None, 4, 9, 6, t_init=torch.zeros(2, 3, device=dev))      # n > N
    with pytest.raises(ValueError):
        native.rslm_draw(None, None, w2d, None, 4, 2, 6, t_init=torch.zeros(3, 3, device=dev))      # t_init of another batch
    with pytest.raises(ValueError):
        native.rslm_draw(None, None, w2d, None, 4, 2, 6)                                            # neither t_init nor the points
    with pytest.raises(ValueError):
        RSLMSolver(dof=6, draws="numpy")


def test_solver_with_native_draws(dev):
    """RSLMSolver end to end on its default (native) draws: reproducible under torch.manual_seed, the returned cost is the
    winner's full-set cost, and as an initialiser it lands LMSolver on the ground truth like the torch draws do."""
    B, N = 4, 64
    pc = {k: v.to(dev) for k, v in make_problem(B, N, seed=50).items()}
    camera = PerspectiveCamera(cam_mats=pc["cam_mats"])
    cost_fun = AdaptiveHuberPnPCost(relative_delta=0.5)
    cost_fun.set_param(pc["x2d"], pc["w2d"])
    rs = RSLMSolver(dof=6, num_points=8, num_proposals=32, num_iter=5)
    assert rs.draws == "native"
    torch.manual_seed(7)
    p1, none, c1 = rs.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
    torch.manual_seed(7)
    p2, _, c2 = rs.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)
    assert none is None and torch.equal(p1, p2) and torch.equal(c1, c2)
    p3, _, _ = rs.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun, rslm_seed=123)
    p4, _, _ = rs.solve(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun, rslm_seed=123)
    assert torch.equal(p3, p4)
    full = evaluate_pnp(pc["x3d"], pc["x2d"], pc["w2d"], p1, camera, cost_fun, out_cost=True)[1]
    assert torch.allclose(full, c1, rtol=1e-5, atol=1e-5)
    gt = pc["pose_gt"]
    hits = {}
    for draws in ("native", "torch"):
        torch.manual_seed(1)
        solver = LMSolver(dof=6, num_iter=10, init_solver=RSLMSolver(dof=6, num_points=8, num_proposals=32, num_iter=5, draws=draws))
        pose = solver(pc["x3d"], pc["x2d"], pc["w2d"], camera, cost_fun)[0]
        hits[draws] = ((pose[:, :3] - gt[:, :3]).norm(dim=-1) < 0.15).float().mean().item()
    assert hits["native"] >= 0.75 and hits["native"] >= hits["torch"] - 0.25, hits
