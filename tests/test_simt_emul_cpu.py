"""The REAL kernel source on the CPU: epro-pnp_b200/csrc/pnp_kernels.cu compiled by g++ against the SIMT emulator
in tests/simt_emul/ (128 fibers per CTA, barriers and warp shuffles emulated), driven through the same C ABI /
epropnp_b200.native / drop-in classes as on the GPU, and held to the SAME assertions as the `-m gpu` parity tests --
the test functions below are the GPU suite's own, re-collected here with a `cuda_device` fixture that installs the
emulated backend.

Scope (see tests/simt_emul/cuda_runtime.h): control flow, shared-memory bookkeeping, reductions, the LM state machine
and the AMIS loop of the shipped kernels, plus the C ABI's argument handling.  Not covered: races, TMA / mbarrier
phases, approximate special-function units, performance -- those are what `-m gpu` is for.
"""
import pytest
import torch

import simt_native
import test_autograd_gpu as _ag
import test_dropin_gpu as _dg
import test_gpu_edges as _ge
import test_gpu_parity as _gp


@pytest.fixture
def cuda_device(monkeypatch):
    return simt_native.install(monkeypatch)


# ---- goldens and oracle comparisons of the GPU parity suite
test_golden_evaluate = _gp.test_golden_evaluate
test_golden_lm_solve = _gp.test_golden_lm_solve
test_golden_amis_from_reference_solution = _gp.test_golden_amis_from_reference_solution
test_golden_fused_lm_amis = _gp.test_golden_fused_lm_amis
test_golden_fused_lm_amis_4dof = _gp.test_golden_fused_lm_amis_4dof

test_fused_against_oracle_north_star_shape = _gp.test_fused_against_oracle_north_star_shape

# ---- edge cases
test_parameter_corners_against_oracle = _ge.test_parameter_corners_against_oracle
test_tiny_point_sets = _ge.test_tiny_point_sets
test_all_points_behind_camera = _ge.test_all_points_behind_camera
test_nan_object_does_not_leak = _ge.test_nan_object_does_not_leak

# ---- the drop-in classes on top of the emulated kernels
test_lmsolver_forward = _dg.test_lmsolver_forward
test_evaluate_pnp_semantics = _dg.test_evaluate_pnp_semantics
test_monte_carlo_forward = _dg.test_monte_carlo_forward
test_layer_forward_and_4dof = _dg.test_layer_forward_and_4dof
# (test_rslm_init_and_force_init_solve is left to the GPU: its ~10^4 tiny LM solves take minutes under emulation)

# ---- backward kernel and the autograd bridge
test_monte_carlo_backward_matches_reference = _ag.test_monte_carlo_backward_matches_reference
test_cost_backward_kernel_against_oracle_autograd = _ag.test_cost_backward_kernel_against_oracle_autograd
test_evaluate_pnp_cost_is_differentiable = _ag.test_evaluate_pnp_cost_is_differentiable


# ------------------------------------------------------------------------------------------------
# Scheduling-order invariance: the emulator runs a CTA's threads one after another between synchronisation points.
# Ascending, descending and randomly permuted orders must give bit-identical results; a difference means one thread
# consumed what another produced without a barrier in between (the class of bug racecheck reports on hardware --
# the round-1 `pad_last_pair` hazard is caught by exactly this test when re-introduced).
def _all_outputs(dev, dof, odd_points):
    from epropnp_b200 import native
    from epropnp_b200.synth import make_problem
    B, N, M, I = 3, (37 if odd_points else 64), 128, 4
    pc = make_problem(B, N, seed=77, dof=dof)
    prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], -50.0, 700.0,
                          native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5))
    p = native.default_params(dof, mc_samples=M, mc_iter=I)
    out = native.lm_amis_fused(prob, pc["pose_init"], p, seed=5, want_cost=True, want_plus=True)
    res = [out[k] for k in ("pose_opt", "pose_cov", "cost", "pose_opt_plus", "pose_samples", "logw")]
    res.append(native.evaluate_cost(prob, out["pose_samples"].transpose(0, 1)[:9].contiguous(), dof, 0.1))
    res.extend(native.evaluate_full(prob, pc["pose_init"], dof, 0.1, 1e-10, True, True, True, True))
    grads = native.cost_backward(prob, dof, 0.1, out["pose_samples"], torch.ones(B, M) / M)
    res.extend(g for g in grads if g is not None)
    return [t.clone() for t in res if t is not None]


@pytest.mark.parametrize("dof,odd_points", [(6, False), (6, True), (4, True)])
def test_results_do_not_depend_on_thread_schedule(monkeypatch, dof, odd_points):
    dev = simt_native.install(monkeypatch)
    lib = simt_native.handle()
    try:
        lib.simt_set_schedule(0, 1)
        base = _all_outputs(dev, dof, odd_points)
        for mode, seed in ((1, 1), (2, 11), (2, 12)):
            lib.simt_set_schedule(mode, seed)
            for a, b in zip(base, _all_outputs(dev, dof, odd_points)):
                assert torch.equal(a, b, ) or (torch.isnan(a) == torch.isnan(b)).all() and torch.equal(
                    torch.nan_to_num(a), torch.nan_to_num(b)), (mode, seed)
    finally:
        lib.simt_set_schedule(0, 1)


@pytest.mark.parametrize("n_chunks,bounded", [(1, False), (3, True), (64, False), (0, False), (0, True)])
def test_host_buffer_entry_point_chunking(cuda_device, monkeypatch, n_chunks, bounded):
    """epnp_lm_amis_fused_host_f32 (host buffers, chunked copy / solve / copy-back pipeline): workspace layout, chunk
    boundaries and per-chunk object offsets of the Philox stream give the device-resident call's results bit for bit.
    n_chunks = 0 cuts at whole waves: a 1-SM "device" with 4 resident CTAs (bounded: 2 SMs) -> chunks of 4 / 8 objects."""
    if n_chunks == 0:
        monkeypatch.setenv("SIMT_EMUL_SMS", "2" if bounded else "1")
    from epropnp_b200 import native
    from epropnp_b200.synth import make_problem
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self, *a, **k: self)       # no CUDA runtime here
    monkeypatch.setattr(torch, "empty", (lambda f: (lambda *a, pin_memory=False, **k: f(*a, **k)))(torch.empty))
    B, N = 11, 24
    pc = make_problem(B, N, seed=8)
    delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
    lb, ub = ((pc["x2d"].amin(1) - 3.0).contiguous(), (pc["x2d"].amax(1) + 3.0).contiguous()) if bounded else (None, None)
    prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], lb, ub, delta)
    p = native.default_params(6, lm_iter=3, mc_samples=16, mc_iter=2)
    ref = native.lm_amis_fused(prob, pc["pose_init"], p, seed=77, obj_offset=5, want_cost=True)
    host = dict(x3d=pc["x3d"], x2d=pc["x2d"], w2d=pc["w2d"], cam_mats=pc["cam_mats"].contiguous(), delta=delta,
                pose_init=pc["pose_init"], lb=lb, ub=ub)
    ws = torch.empty(native.fused_workspace_bytes(B, N, p), dtype=torch.uint8)
    res = native.lm_amis_fused_host(host, p, ws, n_chunks=n_chunks, seed=77, obj_offset=5)
    for k in ("pose_opt", "logw", "pose_samples", "pose_cov", "cost"):
        assert torch.equal(res[k], ref[k]), k
    with pytest.raises(native.NativeError):                                            # workspace one byte short
        native.lm_amis_fused_host(host, p, ws[:-1], n_chunks=n_chunks, seed=77)


# ------------------------------------------------------------------------------------------------
# The size-independent properties the GPU suite checks at B = 4096 (determinism, shard invariance with object
# offsets, TMA vs plain loader, point-permutation invariance, sanity of the weights), at a size the emulator affords.
_BIG = {}


def _small_big():
    from epropnp_b200 import native
    from epropnp_b200.synth import make_problem
    d = {k: v for k, v in make_problem(12, 512, seed=7).items()}
    d["delta"] = native.adaptive_delta(d["x2d"], d["w2d"], 0.5)
    d["prob"] = native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, d["delta"])
    d["params"] = native.default_params(6, mc_samples=512, mc_iter=4)
    d["out"] = native.lm_amis_fused(d["prob"], d["pose_init"], d["params"], seed=1234, want_cost=True)
    return d


@pytest.fixture
def big(monkeypatch):
    simt_native.install(monkeypatch)
    if "d" not in _BIG:
        _BIG["d"] = _small_big()
    return _BIG["d"]


test_full_size_sanity = _gp.test_full_size_sanity
test_full_size_deterministic_and_shard_invariant = _gp.test_full_size_deterministic_and_shard_invariant
test_lm_flip_rate_at_north_star_shape = _gp.test_lm_flip_rate_at_north_star_shape
test_tma_and_plain_loader_agree = _gp.test_tma_and_plain_loader_agree
test_point_permutation_invariance = _gp.test_point_permutation_invariance


def test_long_point_sets_take_the_multi_warp_kernels(cuda_device):
    """N >= 2048 selects the 8-warps-per-object LM kernel and the 512-thread AMIS kernel (split cost sweep): both against
    the fp64 oracle at a size the emulator affords, LM / GN and bounded / unbounded."""
    from oracle import pnp_oracle as orc
    from epropnp_b200 import native
    from epropnp_b200.synth import make_noise, make_problem
    from conftest import err_vs
    B, N, M, I = 2, 2050, 16, 2
    for fast, bounded in ((1, False), (0, True)):
        pc = make_problem(B, N, seed=21 + fast)
        noise = make_noise(B, M, seed=22)
        lb, ub = ((pc["x2d"].amin(1) + 4.0, pc["x2d"].amax(1) - 4.0) if bounded else (None, None))
        delta = native.adaptive_delta(pc["x2d"], pc["w2d"], 0.5)
        prob = native.Problem(pc["x3d"], pc["x2d"], pc["w2d"], pc["cam_mats"], lb, ub, delta)
        p = native.default_params(6, lm_iter=3, fast_mode=fast, mc_samples=M, mc_iter=I)
        out = native.lm_amis_fused(prob, pc["pose_init"], p, noise=noise, want_cost=True)
        d = torch.float64
        cam = orc.Camera(pc["cam_mats"].to(d), 0.1, None if lb is None else lb.to(d), None if ub is None else ub.to(d))
        d64 = orc.adaptive_delta(pc["x2d"].to(d), pc["w2d"].to(d), 0.5)
        pose64, cov64, cost64 = orc.lm_solve(pc["x3d"].to(d), pc["x2d"].to(d), pc["w2d"].to(d), cam, d64, pc["pose_init"].to(d),
                                             orc.LMParams(num_iter=3), fast_mode=bool(fast))
        assert err_vs(out["pose_opt"].numpy(), pose64.numpy()) < 1e-4
        assert err_vs(out["cost"].numpy(), cost64.numpy()) < 1e-4
        S = M // I
        nz = (noise[0].reshape(B, I, S, 3).permute(1, 2, 0, 3).to(d), noise[1].reshape(B, I, S).permute(1, 2, 0).to(d),
              noise[2].reshape(B, I, S, 4).permute(1, 2, 0, 3).to(d))
        r = orc.amis_6dof(pc["x3d"].to(d), pc["x2d"].to(d), pc["w2d"].to(d), cam, d64, out["pose_opt"].to(d), out["pose_cov"].to(d), nz, M, I)
        f = torch.float32                                       # the same loop in fp32: what rounding alone moves
        cam32 = orc.Camera(pc["cam_mats"], 0.1, lb, ub)
        r32 = orc.amis_6dof(pc["x3d"], pc["x2d"], pc["w2d"], cam32, orc.adaptive_delta(pc["x2d"], pc["w2d"], 0.5), out["pose_opt"],
                            out["pose_cov"], tuple(t.to(f) for t in nz), M, I)
        floor_w = err_vs(r32["logw"].numpy(), r["logw"].numpy())
        assert err_vs(out["pose_samples"].transpose(0, 1).numpy(), r["samples"].numpy()) < 2e-4
        assert err_vs(out["logw"].transpose(0, 1).numpy(), r["logw"].numpy()) < max(2e-4, 5 * floor_w)
