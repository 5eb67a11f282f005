"""Property tests (hypothesis) of the scalar code the kernels inline, driven through the host build
(tests/host_emul.cpp): invariances the algorithm must have regardless of size or seed."""
import ctypes

import numpy as np
from hypothesis import HealthCheck, given, settings, strategies as st

from epropnp_b200 import build
from epropnp_b200.capi import EpnpParams
from epropnp_b200.synth import make_problem

EMUL = None


def emul():
    global EMUL
    if EMUL is None:
        EMUL = ctypes.CDLL(build.build_host_emul())
    return EMUL


def fptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def params(dof=6, **kw):
    p = EpnpParams(dof=dof, lm_iter=10, fast_mode=0, z_min=0.1, min_lm_diagonal=1e-6, max_lm_diagonal=1e32,
                   min_relative_decrease=1e-3, initial_radius=30.0, max_radius=1e16, eps=1e-5, huber_eps=1e-10,
                   mc_samples=64, mc_iter=2, amis_eps=1e-5, acg_mle_iter=3, acg_dispersion=1e-3)
    for k, v in kw.items():
        setattr(p, k, v)
    return p


def problem(B, N, seed, dof=6):
    pc = make_problem(B, N, seed=seed, dof=dof)
    a = {k: np.ascontiguousarray(v.numpy(), np.float32) for k, v in pc.items()}
    x2d, w2d = pc["x2d"], pc["w2d"]
    a["delta"] = np.ascontiguousarray((w2d.mean(dim=(-2, -1)) * x2d.var(dim=-2).sum(-1).sqrt() * 0.5).numpy(), np.float32)
    return a


def lm(a, dof=6, **kw):
    B, N = a["x3d"].shape[:2]
    PD = 7 if dof == 6 else 4
    pose = np.zeros((B, PD), np.float32)
    cost = np.zeros(B, np.float32)
    cov = np.zeros((B, dof, dof), np.float32)
    p = params(dof, **kw)
    emul().emul_lm(fptr(a["x3d"]), fptr(a["x2d"]), fptr(a["w2d"]), fptr(a["cam_mats"]), None, None, fptr(a["delta"]),
                   fptr(a["pose_init"]), fptr(pose), fptr(cov), fptr(cost), None, None, B, N, ctypes.byref(p))
    return pose, cost, cov


fast = settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))


@fast
@given(seed=st.integers(0, 10 ** 6), N=st.integers(12, 96), dof=st.sampled_from([4, 6]))
def test_lm_is_invariant_to_point_order(seed, N, dof):
    a = problem(3, N, seed, dof)
    pose, cost, _ = lm(a, dof)
    perm = np.random.default_rng(seed).permutation(N)
    b = dict(a)
    for k in ("x3d", "x2d", "w2d"):
        b[k] = np.ascontiguousarray(a[k][:, perm])
    pose_p, cost_p, _ = lm(b, dof)
    assert np.abs(cost - cost_p).max() <= 2e-5 * np.abs(cost).max() + 1e-6
    assert np.abs(pose - pose_p).max() < 2e-3        # flat directions may flip an accept/reject; cost is the invariant


@fast
@given(seed=st.integers(0, 10 ** 6), N=st.integers(12, 64))
def test_centering_the_points_does_not_change_the_solution(seed, N):
    """pnp_normalize -> solve -> pnp_denormalize == plain solve (common.py:103-136): shifting x3d by its mean and
    the translation by R * mean is an exact reparameterisation of the same cost."""
    a = problem(2, N, seed)
    pose, cost, _ = lm(a)
    off = a["x3d"].mean(axis=1, keepdims=True)
    b = dict(a)
    b["x3d"] = np.ascontiguousarray(a["x3d"] - off)

    def rot(q):
        w, x, y, z = q
        return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]], np.float64)
    p0 = a["pose_init"].copy()
    for i in range(2):
        p0[i, :3] += (rot(p0[i, 3:].astype(np.float64)) @ off[i, 0].astype(np.float64)).astype(np.float32)
    b["pose_init"] = p0
    pose_n, cost_n, _ = lm(b)
    back = pose_n.copy()
    for i in range(2):
        back[i, :3] -= (rot(pose_n[i, 3:].astype(np.float64)) @ off[i, 0].astype(np.float64)).astype(np.float32)
    assert np.abs(cost - cost_n).max() <= 5e-5 * np.abs(cost).max() + 1e-6
    assert np.abs(back - pose).max() < 2e-3


@fast
@given(seed=st.integers(0, 10 ** 6), N=st.integers(8, 40))
def test_reverse_mode_matches_finite_differences(seed, N):
    a = problem(2, N, seed)
    B, P = 2, 3
    rng = np.random.default_rng(seed)
    poses = np.repeat(a["pose_gt"][:, None, :], P, 1).astype(np.float32)
    poses[..., :3] += 0.05 * rng.standard_normal((B, P, 3)).astype(np.float32)
    up = rng.standard_normal((B, P)).astype(np.float32)
    delta = np.full(B, 1e9, np.float32)           # quadratic everywhere: smooth, finite differences are clean
    g3 = np.zeros((B, N, 3), np.float32); g2 = np.zeros((B, N, 2), np.float32)
    gw = np.zeros((B, N, 2), np.float32); gd = np.zeros(B, np.float32)
    e = emul()
    e.emul_cost_backward(fptr(a["x3d"]), fptr(a["x2d"]), fptr(a["w2d"]), fptr(a["cam_mats"]), None, None, fptr(delta),
                         fptr(poses), fptr(up), fptr(g3), fptr(g2), fptr(gw), fptr(gd), P, B, N, 6, ctypes.c_float(0.1))

    def total(x3d, x2d, w2d):
        c = np.zeros((P, B), np.float32)
        pz = np.ascontiguousarray(np.transpose(poses, (1, 0, 2)))
        e.emul_cost(fptr(x3d), fptr(x2d), fptr(w2d), fptr(a["cam_mats"]), None, None, fptr(delta), fptr(pz), fptr(c), P, B,
                    N, 6, ctypes.c_float(0.1))
        return float((c.astype(np.float64) * up.T.astype(np.float64)).sum())
    for name, g, h in (("x3d", g3, 2e-3), ("x2d", g2, 0.2), ("w2d", gw, 2e-3)):
        v = rng.standard_normal(a[name].shape).astype(np.float32)
        args_p = {k: a[k] for k in ("x3d", "x2d", "w2d")}
        args_m = dict(args_p)
        args_p[name] = np.ascontiguousarray(a[name] + h * v)
        args_m[name] = np.ascontiguousarray(a[name] - h * v)
        fd = (total(**args_p) - total(**args_m)) / (2 * h)
        an = float((g.astype(np.float64) * v.astype(np.float64)).sum())
        assert abs(fd - an) <= 3e-2 * max(abs(an), 1e-3) + 1e-2, (name, fd, an)


@fast
@given(seed=st.integers(0, 10 ** 6))
def test_amis_weights_are_a_valid_importance_sample(seed):
    """Self-normalised weights: finite, and the proposal adapts (ESS of the last iteration's samples is not worse
    than the first's by more than noise) -- on the emulated production path (Philox draws)."""
    a = problem(2, 48, seed)
    pose, cost, cov = lm(a)
    B, M, I = 2, 256, 4
    p = params(mc_samples=M, mc_iter=I)
    smp = np.zeros((B, M, 7), np.float32); logw = np.zeros((B, M), np.float32)
    emul().emul_amis6(fptr(a["x3d"]), fptr(a["x2d"]), fptr(a["w2d"]), fptr(a["cam_mats"]), None, None, fptr(a["delta"]),
                      fptr(pose), fptr(cov), None, None, None, ctypes.c_uint64(seed), ctypes.c_uint32(0), fptr(smp),
                      fptr(logw), None, B, 48, ctypes.byref(p))
    assert np.isfinite(logw).all() and np.isfinite(smp).all()
    assert np.abs(np.linalg.norm(smp[..., 3:], axis=-1) - 1).max() < 1e-5
    w = np.exp(logw - logw.max(1, keepdims=True))
    w /= w.sum(1, keepdims=True)
    assert (1.0 / (w ** 2).sum(1) > 8).all()
