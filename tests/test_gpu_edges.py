"""Edge cases of the native path (GPU): degenerate sizes and parameters, clamped geometry, NaN isolation,
batch sizes that do not divide the persistent grid, the largest resident N.  Compared with the CPU oracle where
the result is well defined, otherwise checked for the reference's documented behaviour (finite / unchanged)."""
import pytest
import torch

from conftest import assert_lm_parity, err_vs
from epropnp_b200 import native
from epropnp_b200.synth import make_noise, make_problem

pytestmark = pytest.mark.gpu


def _prob(pc, dev, delta=None, rel=0.5):
    d = {k: v.to(dev) for k, v in pc.items()}
    if delta is None:
        delta = native.adaptive_delta(d["x2d"], d["w2d"], rel)
    return native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, delta), d


def _oracle(pc, noise, M, I, dtype=torch.float64, **kw):
    from oracle import pnp_oracle as orc
    t = lambda k: pc[k].to(dtype)
    B, S = pc["x3d"].shape[0], M // I
    n3, c2, n4 = noise
    nz = (n3.reshape(B, I, S, 3).permute(1, 2, 0, 3).to(dtype), c2.reshape(B, I, S).permute(1, 2, 0).to(dtype),
          n4.reshape(B, I, S, 4).permute(1, 2, 0, 3).to(dtype))
    cam = orc.Camera(t("cam_mats"), 0.1)
    delta = orc.adaptive_delta(t("x2d"), t("w2d"), 0.5)
    pose, cov, cost = orc.lm_solve(t("x3d"), t("x2d"), t("w2d"), cam, delta, t("pose_init"),
                                   orc.LMParams(num_iter=kw.get("lm_iter", 10)))
    r = orc.amis_6dof(t("x3d"), t("x2d"), t("w2d"), cam, delta, pose, cov, nz, M, I,
                      acg_mle_iter=kw.get("acg_mle_iter", 3))
    r.update(pose_opt=pose, pose_cov=cov, lm_cost=cost)
    return r


@pytest.mark.parametrize("M,I,acg,lm_iter", [(128, 1, 3, 10), (64, 2, 1, 10), (96, 3, 0, 10), (128, 4, 3, 0),
                                             (1024, 8, 2, 4)])
def test_parameter_corners_against_oracle(cuda_device, M, I, acg, lm_iter):
    B, N = 6, 96
    pc = make_problem(B, N, seed=M + I)
    noise = make_noise(B, M, seed=3)
    ref = _oracle(pc, noise, M, I, acg_mle_iter=acg, lm_iter=lm_iter)
    ref32 = _oracle(pc, noise, M, I, dtype=torch.float32, acg_mle_iter=acg, lm_iter=lm_iter)
    prob, d = _prob(pc, cuda_device)
    p = native.default_params(6, mc_samples=M, mc_iter=I, acg_mle_iter=acg, lm_iter=lm_iter)
    out = native.lm_amis_fused(prob, d["pose_init"], p, noise=tuple(t.to(cuda_device) for t in noise), want_cost=True)
    assert_lm_parity(out["pose_opt"].cpu().numpy(), out["cost"].cpu().numpy(), ref["pose_opt"].numpy(),
                     ref["lm_cost"].numpy(), 1e-4, what="corner pose")
    floor = err_vs(ref32["logw"], ref["logw"])
    assert err_vs(out["logw"].transpose(0, 1).cpu(), ref["logw"]) < max(1e-4, 5 * floor)
    if lm_iter == 0:
        assert torch.equal(out["pose_opt"], d["pose_init"])


@pytest.mark.parametrize("N", [1, 2, 3, 5])
def test_tiny_point_sets(cuda_device, N):
    """Fewer equations than unknowns: J^T J is singular, only the damping / eps keeps the solves defined."""
    from oracle import pnp_oracle as orc
    B = 8
    pc = make_problem(B, N, seed=40 + N)
    prob, d = _prob(pc, cuda_device, delta=1.0)
    p = native.default_params(6, mc_samples=64, mc_iter=2)
    out = native.lm_amis_fused(prob, d["pose_init"], p, seed=1, want_cost=True)
    assert torch.isfinite(out["pose_opt"]).all() and torch.isfinite(out["cost"]).all()
    assert torch.isfinite(out["logw"]).all() and torch.isfinite(out["pose_samples"]).all()
    d64 = torch.float64
    cam = orc.Camera(pc["cam_mats"].to(d64), 0.1)
    _, _, cost64 = orc.lm_solve(pc["x3d"].to(d64), pc["x2d"].to(d64), pc["w2d"].to(d64), cam, 1.0, pc["pose_init"].to(d64))
    c0 = orc.evaluate(pc["x3d"].to(d64), pc["x2d"].to(d64), pc["w2d"].to(d64), pc["pose_init"].to(d64), cam, 1.0)["cost"]
    # the solve must not do worse than where it started and should reach the oracle's cost level
    assert (out["cost"].cpu().double() <= c0 * (1 + 1e-5) + 1e-6).all()
    assert (out["cost"].cpu().double() <= cost64 * 1.05 + 1e-3).all()


def test_all_points_behind_camera(cuda_device):
    """Every projection sits on the z clamp: clipped Jacobians are zero, the pose must stay put
    (camera.py:100-105), covariance = I / eps, AMIS still returns finite weights."""
    B, N = 4, 64
    pc = make_problem(B, N, seed=5)
    pc["pose_init"][:, 2] = -6.0
    prob, d = _prob(pc, cuda_device)
    p = native.default_params(6, mc_samples=128, mc_iter=4)
    out = native.lm_amis_fused(prob, d["pose_init"], p, seed=2, want_cost=True)
    assert torch.allclose(out["pose_opt"], d["pose_init"], atol=1e-6)
    eye = torch.eye(6, device=cuda_device) / 1e-5
    assert torch.allclose(out["pose_cov"], eye.expand(B, 6, 6), rtol=1e-3)
    assert torch.isfinite(out["logw"]).all()


def test_nan_object_does_not_leak(cuda_device):
    B, N = 12, 64
    pc = make_problem(B, N, seed=6)
    prob, d = _prob(pc, cuda_device)
    p = native.default_params(6, mc_samples=128, mc_iter=4)
    clean = native.lm_amis_fused(prob, d["pose_init"], p, seed=3, want_cost=True)
    bad = {k: v.clone() for k, v in pc.items()}
    bad["x2d"][5, 7, 0] = float("nan")
    prob_b, db = _prob(bad, cuda_device)
    prob_b.delta = prob.delta.clone()
    out = native.lm_amis_fused(prob_b, db["pose_init"], p, seed=3, want_cost=True)
    torch.cuda.synchronize()                                   # no hang
    keep = [i for i in range(B) if i != 5]
    for k in ("pose_opt", "logw", "pose_samples", "cost"):
        assert torch.equal(out[k][keep], clean[k][keep]), k
    assert torch.equal(out["pose_opt"][5], db["pose_init"][5])   # every step rejected (NaN cost) -> pose unchanged
    assert not torch.isfinite(out["logw"][5]).any()


@pytest.mark.parametrize("B", [1, 3, 593, 1187])
def test_batch_sizes_off_the_grid(cuda_device, B):
    pc = make_problem(B, 32, seed=B)
    prob, d = _prob(pc, cuda_device)
    p = native.default_params(6, mc_samples=64, mc_iter=4)
    out = native.lm_amis_fused(prob, d["pose_init"], p, seed=9, want_cost=True)
    # object b of the batch == the same object solved alone with its global index
    for b in {0, B // 2, B - 1}:
        sub = native.Problem(d["x3d"][b:b + 1], d["x2d"][b:b + 1], d["w2d"][b:b + 1], d["cam_mats"][b:b + 1], None, None,
                             prob.delta[b:b + 1])
        one = native.lm_amis_fused(sub, d["pose_init"][b:b + 1], p, seed=9, obj_offset=b, want_cost=True)
        assert torch.equal(one["logw"][0], out["logw"][b]) and torch.equal(one["pose_opt"][0], out["pose_opt"][b])


def test_largest_resident_point_set(cuda_device):
    """N = epnp_max_points(6, 512, 4): one object fills the 227 KB of an SM; N + 4 is refused with an error."""
    nmax = native.capi.lib().epnp_max_points(6, 512, 4)
    assert nmax >= 4096
    pc = make_problem(3, nmax, seed=8)
    prob, d = _prob(pc, cuda_device)
    p = native.default_params(6, mc_samples=512, mc_iter=4)
    out = native.lm_amis_fused(prob, d["pose_init"], p, seed=1, want_cost=True)
    gt = d["pose_gt"]
    assert (out["pose_opt"][:, :3] - gt[:, :3]).norm(dim=-1).max() < 0.02
    assert torch.isfinite(out["logw"]).all()
    lm_only_max = native.capi.lib().epnp_max_points(6, 0, 0)
    assert lm_only_max > nmax
