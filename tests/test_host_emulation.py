"""CPU check of the scalar code the kernels inline (epro-pnp_b200/csrc/pnp_math.cuh), compiled with
g++ and driven serially by tests/host_emul.cpp, against the reference's golden vectors.
This is host-logic coverage for `-m "not gpu"`; the CUDA path itself is tested in test_gpu_parity.py."""
import ctypes

import numpy as np
import pytest
import torch

from conftest import assert_lm_parity, err_vs, golden_names, load_golden
from epropnp_b200 import build
from epropnp_b200.capi import EpnpParams


@pytest.fixture(scope="module")
def emul():
    return ctypes.CDLL(build.build_host_emul())


def fptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def params_for(g, **kw):
    p = EpnpParams(dof=int(g["dof"]), lm_iter=int(g["lm_iter"]), fast_mode=int(g["fast_mode"]),
                   z_min=float(g["z_min"]), min_lm_diagonal=1e-6, max_lm_diagonal=1e32,
                   min_relative_decrease=1e-3, initial_radius=30.0, max_radius=1e16, eps=1e-5,
                   huber_eps=1e-10, mc_samples=512, mc_iter=4, amis_eps=1e-5, acg_mle_iter=3,
                   acg_dispersion=1e-3)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def inputs(g):
    B, N = int(g["B"]), int(g["N"])
    f = lambda k: np.ascontiguousarray(g[k], dtype=np.float32)
    kind = int(g["bounds"])
    if kind == 0:
        lb = ub = None
    elif kind == 1:
        lb = np.full((B, 2), float(g["lb"]), np.float32)
        ub = np.full((B, 2), float(g["ub"]), np.float32)
    else:
        lb, ub = f("lb"), f("ub")
    return B, N, f("x3d"), f("x2d"), f("w2d"), f("cam_mats"), lb, ub, f("delta"), f("pose_init")


@pytest.mark.parametrize("name", golden_names())
def test_point_math(emul, name):
    g = load_golden(name)
    B, N, x3d, x2d, w2d, cam, lb, ub, delta, pose = inputs(g)
    dof = int(g["dof"])
    clip = 0 if int(g["fast_mode"]) else 1
    res = np.zeros((B, 2 * N), np.float32)
    jac = np.zeros((B, 2 * N, dof), np.float32)
    cost = np.zeros(B, np.float32)
    emul.emul_residual_jac(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose),
                           fptr(res), fptr(jac), fptr(cost), clip, B, N, dof, ctypes.c_float(float(g["z_min"])),
                           ctypes.c_float(1e-10))
    # scale-relative against the fp64 run of the reference: fp32 rounding only
    assert err_vs(res, g["ref64_eval_residual"]) < 2e-4
    assert err_vs(jac, g["ref64_eval_jac"]) < 2e-5
    assert err_vs(cost, g["ref64_eval_cost"]) < 2e-5
    # normal equations = J^T J, J^T r, cost of the golden Jacobian
    NV = dof * (dof + 1) // 2 + dof + 1
    J, r = g["ref64_eval_jac"], g["ref64_eval_residual"]
    JtJ = np.einsum("bnd,bne->bde", J, J)
    iu = np.triu_indices(dof)
    want = np.concatenate([JtJ[:, iu[0], iu[1]], np.einsum("bnd,bn->bd", J, r), g["ref64_eval_cost"][:, None]], 1)
    # scalar form (the CTA-per-object kernels) and the row-packed form (the warp-per-object LM kernel): same bounds
    for fn in (emul.emul_normal_eq, emul.emul_normal_eq_rows):
        ne = np.zeros((B, NV), np.float32)
        fn(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose),
           fptr(ne), clip, B, N, dof, ctypes.c_float(float(g["z_min"])), ctypes.c_float(1e-10))
        for b in range(B):
            assert err_vs(ne[b, :-dof - 1], want[b, :-dof - 1]) < 5e-5
            assert err_vs(ne[b, -dof - 1:-1], want[b, -dof - 1:-1]) < 5e-4    # J^T r: cancellation near the optimum
        assert err_vs(ne[:, -1], g["ref64_eval_cost"]) < 2e-5
    # cost of stacked poses (pre-multiplied projection path)
    poses = np.ascontiguousarray(g["eval_poses"], np.float32)
    S = poses.shape[0]
    for fn in (emul.emul_cost,):
        cm = np.zeros((S, B), np.float32)
        fn(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(poses), fptr(cm),
           S, B, N, dof, ctypes.c_float(float(g["z_min"])))
        assert err_vs(cm, g["ref64_eval_cost_multi"]) < 2e-5


@pytest.mark.parametrize("name", golden_names())
def test_lm_state_machine(emul, name):
    g = load_golden(name)
    B, N, x3d, x2d, w2d, cam, lb, ub, delta, pose0 = inputs(g)
    dof = int(g["dof"])
    PD = 7 if dof == 6 else 4
    p = params_for(g)
    pose = np.zeros((B, PD), np.float32)
    cov = np.zeros((B, dof, dof), np.float32)
    cost = np.zeros(B, np.float32)
    plus = np.zeros((B, PD), np.float32)
    cinit = np.zeros(B, np.float32)
    emul.emul_lm(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose0), fptr(pose),
                 fptr(cov), fptr(cost), fptr(plus), fptr(cinit), B, N, ctypes.byref(p))
    floor = err_vs(g["ref32_lm_pose"], g["ref64_lm_pose"])
    tol = max(1e-4, 3 * floor)
    assert_lm_parity(pose, cost, g["ref64_lm_pose"], g["ref64_lm_cost"], tol, what=name + " vs ref64")
    assert_lm_parity(pose, cost, g["ref32_lm_pose"], g["ref64_lm_cost"], tol, what=name + " vs ref32")
    assert err_vs(cost, g["ref64_lm_cost"]) < max(1e-4, 3 * err_vs(g["ref32_lm_cost"], g["ref64_lm_cost"]))
    assert err_vs(cov, g["ref64_lm_cov"]) < max(2e-3, 3 * err_vs(g["ref32_lm_cov"], g["ref64_lm_cov"]))
    assert_lm_parity(plus, cost, g["ref64_lm_pose_plus"], g["ref64_lm_cost"], tol, what=name + " pose_plus")
    assert err_vs(cinit, g["ref64_eval_cost"]) < 2e-5


def _noise_object_major(g):
    n3 = np.ascontiguousarray(np.transpose(g["noise_normal"], (2, 0, 1, 3)).reshape(int(g["B"]), -1, 3), np.float32)
    c2 = np.ascontiguousarray(np.transpose(g["noise_chi2"], (2, 0, 1)).reshape(int(g["B"]), -1), np.float32)
    n4 = np.ascontiguousarray(np.transpose(g["noise_rot"], (2, 0, 1, 3)).reshape(int(g["B"]), -1, 4), np.float32)
    return n3, c2, n4


@pytest.mark.parametrize("name", golden_names("mc6"))
@pytest.mark.parametrize("start", ["golden_lm", "own_lm"])
def test_amis_formulas(emul, name, start):
    g = load_golden(name)
    B, N, x3d, x2d, w2d, cam, lb, ub, delta, pose0 = inputs(g)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    p = params_for(g, mc_samples=M, mc_iter=I)
    if start == "golden_lm":
        pose = np.ascontiguousarray(g["ref32_lm_pose"], np.float32)
        cov = np.ascontiguousarray(g["ref32_lm_cov"], np.float32)
    else:
        pose = np.zeros((B, 7), np.float32)
        cov = np.zeros((B, 6, 6), np.float32)
        emul.emul_lm(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose0), fptr(pose),
                     fptr(cov), None, None, None, B, N, ctypes.byref(p))
    n3, c2, n4 = _noise_object_major(g)
    smp = np.zeros((B, M, 7), np.float32)
    logw = np.zeros((B, M), np.float32)
    props = np.zeros((B, I, 19), np.float32)
    emul.emul_amis6(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose), fptr(cov),
                    fptr(n3), fptr(c2), fptr(n4), ctypes.c_uint64(0), ctypes.c_uint32(0), fptr(smp), fptr(logw),
                    fptr(props), B, N, ctypes.byref(p))
    smp_r = np.transpose(smp, (1, 0, 2))
    logw_r = logw.T
    floor_s = err_vs(g["ref32_mc_samples"], g["ref64_mc_samples"])
    floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
    # max-norm (dominated by a few heavy-tail samples) within 5x of the reference's own fp32 floor,
    # bulk of the distribution (median, 99th percentile) within 3x
    assert err_vs(smp_r, g["ref64_mc_samples"]) < max(1e-4, 5 * floor_s)
    assert err_vs(logw_r, g["ref64_mc_logw"]) < max(1e-4, 5 * floor_w)
    mine = np.abs(logw_r - g["ref64_mc_logw"])
    ref = np.abs(g["ref32_mc_logw"] - g["ref64_mc_logw"])
    for q in (50, 99):
        assert np.percentile(mine, q) < 3 * np.percentile(ref, q) + 1e-5
    # proposals of every AMIS iteration
    mode = np.transpose(props[:, :, :3], (1, 0, 2))
    assert err_vs(mode, g["ref64_mc_trans_mode"]) < 1e-4
    lt = props[:, :, 3:9]
    ref_lt = g["ref64_mc_trans_cov_tril"]
    lt_full = np.stack([lt[..., 0], lt[..., 1], lt[..., 2], lt[..., 3], lt[..., 4], lt[..., 5]], -1)
    ref_pack = np.stack([ref_lt[..., 0, 0], ref_lt[..., 1, 0], ref_lt[..., 1, 1], ref_lt[..., 2, 0],
                         ref_lt[..., 2, 1], ref_lt[..., 2, 2]], -1)
    fl = err_vs(g["ref32_mc_trans_cov_tril"], ref_lt)
    assert err_vs(np.transpose(lt_full, (1, 0, 2)), ref_pack) < max(1e-3, 3 * fl)
    ref_lr = g["ref64_mc_rot_cov_tril"]
    rp = np.stack([ref_lr[..., i, j] for i in range(4) for j in range(i + 1)], -1)
    fr = err_vs(g["ref32_mc_rot_cov_tril"], ref_lr)
    assert err_vs(np.transpose(props[:, :, 9:], (1, 0, 2)), rp) < max(1e-3, 3 * fr)


def test_amis4_formulas(emul):
    g = load_golden("mc4_basic")
    B, N, x3d, x2d, w2d, cam, lb, ub, delta, pose0 = inputs(g)
    M, I = int(g["mc_samples_total"]), int(g["mc_iters"])
    p = params_for(g, mc_samples=M, mc_iter=I)
    pose = np.ascontiguousarray(g["ref32_lm_pose"], np.float32)
    cov = np.ascontiguousarray(g["ref32_lm_cov"], np.float32)
    n3 = np.ascontiguousarray(np.transpose(g["noise_normal"], (2, 0, 1, 3)).reshape(B, -1, 3), np.float32)
    c2 = np.ascontiguousarray(np.transpose(g["noise_chi2"], (2, 0, 1)).reshape(B, -1), np.float32)
    yaw = np.ascontiguousarray(np.transpose(g["yaw_samples"], (2, 0, 1)).reshape(B, -1), np.float32)
    smp = np.zeros((B, M, 4), np.float32)
    logw = np.zeros((B, M), np.float32)
    props = np.zeros((B, I, 19), np.float32)
    emul.emul_amis4(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(pose), fptr(cov),
                    fptr(n3), fptr(c2), fptr(yaw), ctypes.c_uint64(0), ctypes.c_uint32(0), fptr(smp), fptr(logw),
                    fptr(props), B, N, ctypes.byref(p))
    floor_w = err_vs(g["ref32_mc_logw"], g["ref64_mc_logw"])
    # compare with the fp32 reference run (the injected yaw draws are that run's)
    assert err_vs(logw.T, g["ref32_mc_logw"]) < max(1e-4, 5 * floor_w)
    assert err_vs(np.transpose(smp, (1, 0, 2)), g["ref32_mc_samples"]) < 1e-4
    assert err_vs(np.transpose(props[:, :, :3], (1, 0, 2)), g["ref32_mc_trans_mode"]) < 1e-4
    assert err_vs(props[:, :, 9].T, g["ref32_mc_rot_mode"][..., 0]) < 1e-4
    assert err_vs(props[:, :, 10].T, g["ref32_mc_rot_kappa"][..., 0]) < 2e-3


@pytest.mark.parametrize("name", ["lm6_basic", "lm6_bounds", "lm6_scalar_bounds_fixed_delta", "lm4_basic"])
def test_cost_backward_math(emul, name):
    """Reverse mode of the cost (what the native backward kernel inlines) against torch autograd of the
    pinned oracle in float64: d sum_p g_p cost(pose_p) / d(x3d, x2d, w2d, delta)."""
    from oracle import pnp_oracle as orc
    from conftest import golden_bounds
    g = load_golden(name)
    B, N, x3d, x2d, w2d, cam, lb, ub, delta, pose0 = inputs(g)
    dof = int(g["dof"])
    poses = np.ascontiguousarray(np.transpose(g["eval_poses"], (1, 0, 2)), np.float32)       # (B, P, D)
    P = poses.shape[1]
    rng = np.random.default_rng(5)
    up = rng.standard_normal((B, P)).astype(np.float32)
    gx3d = np.zeros((B, N, 3), np.float32); gx2d = np.zeros((B, N, 2), np.float32)
    gw2d = np.zeros((B, N, 2), np.float32); gdel = np.zeros(B, np.float32)
    emul.emul_cost_backward(fptr(x3d), fptr(x2d), fptr(w2d), fptr(cam), fptr(lb), fptr(ub), fptr(delta), fptr(poses),
                            fptr(up), fptr(gx3d), fptr(gx2d), fptr(gw2d), fptr(gdel), P, B, N, dof,
                            ctypes.c_float(float(g["z_min"])))
    d = torch.float64
    t3, t2, tw = (torch.from_numpy(a).to(d).requires_grad_(True) for a in (x3d, x2d, w2d))
    td = torch.from_numpy(delta).to(d).requires_grad_(True)
    lb_o, ub_o = golden_bounds(g, d)
    ocam = orc.Camera(torch.from_numpy(cam).to(d), float(g["z_min"]), lb_o, ub_o)
    cost = orc.evaluate(t3, t2, tw, torch.from_numpy(g["eval_poses"]).to(d), ocam, td)["cost"]        # (P, B)
    (cost * torch.from_numpy(up.T.copy()).to(d)).sum().backward()
    assert err_vs(gx3d, t3.grad.numpy()) < 2e-4
    assert err_vs(gx2d, t2.grad.numpy()) < 2e-4
    assert err_vs(gw2d, tw.grad.numpy()) < 2e-4
    assert err_vs(gdel, td.grad.numpy()) < 2e-4


def test_production_yaw_sampler(emul):
    """Best-Fisher von Mises + 25 % uniform mixture: circular moments of the draws."""
    n, S = 200000, 128
    for kappa in (0.5, 4.0, 50.0, 2000.0):
        out = np.zeros(n, np.float32)
        emul.emul_yaw(ctypes.c_uint64(9), ctypes.c_uint32(3), n, S, ctypes.c_float(0.7), ctypes.c_float(kappa), fptr(out))
        assert np.abs(out).max() <= np.pi + 1e-5
        is_uniform = (np.arange(n) % S) < 32
        u, v = out[is_uniform].astype(np.float64), out[~is_uniform].astype(np.float64)
        assert abs(np.cos(u).mean()) < 0.02 and abs(np.sin(u).mean()) < 0.02          # flat on the circle
        from scipy.special import i0e, i1e
        a1 = i1e(kappa) / i0e(kappa)                                                  # mean resultant length
        assert abs(np.cos(v - 0.7).mean() - a1) < 0.01
        assert abs(np.sin(v - 0.7).mean()) < 0.01


def test_production_rng_statistics(emul):
    """Philox + Box-Muller base noise: moments of the three noise families."""
    n = 200000
    out = np.zeros((n, 8), np.float32)
    emul.emul_base_noise(ctypes.c_uint64(1234), ctypes.c_uint32(7), n, fptr(out))
    z = np.concatenate([out[:, :3], out[:, 4:]], 1).astype(np.float64)
    assert np.abs(z.mean(0)).max() < 0.01
    assert np.abs(z.var(0) - 1).max() < 0.02
    assert np.abs(np.corrcoef(z.T) - np.eye(7)).max() < 0.01
    assert abs((z ** 4).mean() - 3.0) < 0.1
    chi = out[:, 3].astype(np.float64)
    assert abs(chi.mean() - 3.0) < 0.03 and abs(chi.var() - 6.0) < 0.15
    # different objects / seeds decorrelate, same key reproduces
    out2 = np.zeros((1000, 8), np.float32)
    emul.emul_base_noise(ctypes.c_uint64(1234), ctypes.c_uint32(8), 1000, fptr(out2))
    assert abs(np.corrcoef(out[:1000, 0], out2[:, 0])[0, 1]) < 0.12
    out3 = np.zeros((1000, 8), np.float32)
    emul.emul_base_noise(ctypes.c_uint64(1234), ctypes.c_uint32(7), 1000, fptr(out3))
    assert np.array_equal(out3, out[:1000])


def test_refit_fallbacks_and_regular_factors(emul):
    """refit_finish6 / refit_finish4 (the tail of estimate_params, epropnp.py:232-260, 317-342): a translation covariance
    that is not positive definite gets the identity (6DoF) / diag(1, 1, 4) (4DoF) like cholesky_wrapper (:16-33), a
    scatter matrix that is not positive definite gets the identity; regular inputs give the Cholesky factors."""
    mean = np.array([0.1, -0.2, 5.0], np.float32)
    bad_tc = np.array([1.0, 0.0, 0.0, -1.0, 0.0, 1.0], np.float32)              # diag(1, -1, 1), packed upper
    A = np.array([[2.0, 0.3, -0.1], [0.3, 1.5, 0.2], [-0.1, 0.2, 0.7]])
    good_tc = np.array([A[0, 0], A[0, 1], A[0, 2], A[1, 1], A[1, 2], A[2, 2]], np.float32)
    S = np.array([[0.4, 0.05, 0.0, 0.02], [0.05, 0.3, 0.01, 0.0], [0.0, 0.01, 0.2, 0.03], [0.02, 0.0, 0.03, 0.1]])
    good_lam = np.array([S[i, j] for i in range(4) for j in range(i, 4)], np.float32)
    bad_lam = good_lam.copy()
    bad_lam[4] = -0.3                                                         # a negative diagonal entry
    out = np.zeros(19, np.float32)
    tril3 = lambda L: np.array([L[0, 0], L[1, 0], L[1, 1], L[2, 0], L[2, 1], L[2, 2]])
    # 6DoF
    emul.emul_refit_finish(6, fptr(mean), fptr(bad_tc), fptr(good_lam), fptr(out))
    assert np.array_equal(out[3:9], np.array([1, 0, 1, 0, 0, 1], np.float32)) and np.array_equal(out[:3], mean)
    emul.emul_refit_finish(6, fptr(mean), fptr(good_tc), fptr(bad_lam), fptr(out))
    assert np.array_equal(out[9:19], np.array([1, 0, 1, 0, 0, 1, 0, 0, 0, 1], np.float32))
    assert np.abs(out[3:9] - tril3(np.linalg.cholesky(A))).max() < 1e-6
    emul.emul_refit_finish(6, fptr(mean), fptr(good_tc), fptr(good_lam), fptr(out))
    Sd = S + np.linalg.det(S) ** 0.25 * 1e-3 * np.eye(4)
    L4 = np.linalg.cholesky(Sd)
    assert np.abs(out[9:19] - np.array([L4[i, j] for i in range(4) for j in range(i + 1)])).max() < 1e-6
    # 4DoF
    sc = np.array([0.3, 0.4], np.float32)
    emul.emul_refit_finish(4, fptr(mean), fptr(bad_tc), fptr(sc), fptr(out))
    assert np.array_equal(out[3:9], np.array([1, 0, 1, 0, 0, 4], np.float32))
    assert abs(out[9] - np.arctan2(0.3, 0.4)) < 1e-6 and abs(out[10] - 0.33 * 0.5 * (2 - 0.25) / 0.75) < 1e-5
    emul.emul_refit_finish(4, fptr(mean), fptr(good_tc), fptr(sc), fptr(out))
    assert np.abs(out[3:9] - tril3(np.linalg.cholesky(A))).max() < 1e-6


@pytest.mark.parametrize("kappa", [1e-7, 1e-5, 1e-4, 1e-3, 3e-2, 0.5, 4.0])
def test_von_mises_sampler_statistics(emul, kappa):
    """draw_yaw (production 4DoF sampler, Best-Fisher rejection): below kappa ~ 1e-3 the textbook constants cancel in
    fp32 (rho = 0, rr = inf: every draw collapsed onto the mode before the fix); the cancellation-free form must give a
    von Mises sample at every concentration -- mean resultant length A(kappa) = I1 / I0, direction = mode."""
    from scipy.special import i0e, i1e
    n, S, mode = 40000, 40000, 0.7
    out = np.zeros(n, np.float32)
    emul.emul_yaw(ctypes.c_uint64(11), ctypes.c_uint32(3), n, S, ctypes.c_float(mode), ctypes.c_float(kappa), fptr(out))
    vm = out[int(np.floor(0.25 * S + 0.5)):].astype(np.float64)   # the first quarter of an iteration is uniform
    assert np.isfinite(vm).all() and np.abs(vm).max() <= np.pi + 1e-5
    assert np.unique(vm).size > 0.9 * vm.size                      # not collapsed onto the mode
    C, Sn = np.cos(vm - mode).mean(), np.sin(vm - mode).mean()
    want = i1e(kappa) / i0e(kappa)
    tol = 4.0 / np.sqrt(vm.size)
    assert abs(C - want) < tol and abs(Sn) < tol, (kappa, C, want, Sn)
