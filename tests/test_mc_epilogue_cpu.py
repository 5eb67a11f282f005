"""Row f4 -- the step after monte_carlo_forward: MonteCarloPoseLoss / softmax weights / Monte-Carlo score.

  * the torch composite (the default) against tests/golden/epilogue/mc_loss.npz, which oracle/make_golden_mc_loss.py records by
    RUNNING the reference's own loss class (6DoF flavour; loss, gradients, EMA norm factor, train and eval mode);
  * the native epilogue kernels (epnp_mc_epilogue_f32 / epnp_mc_lse_backward_f32), executed from the real kernel source
    under the CPU SIMT emulator, against the same vectors, including torch.logsumexp's conventions at NaN / -inf / +inf.
The GPU twin (tests/test_mc_epilogue_gpu.py) re-runs the native half on hardware."""
import numpy as np
import pytest
import torch

import simt_native
from conftest import load_golden
from epropnp import monte_carlo_pose_loss as mcl
from epropnp_b200 import native

G = load_golden("epilogue/mc_loss")
FINITE = [b for b in range(G["logw"].shape[1]) if b != 5]        # object 5 holds a NaN log-weight


@pytest.fixture
def dev(monkeypatch):
    return simt_native.install(monkeypatch)


@pytest.fixture
def native_on(monkeypatch):
    monkeypatch.setattr(mcl, "_use_native", lambda t: True)


def layer_view(a, dev, dtype=torch.float32):
    """(M, B, ...) golden array -> the layer's output format: a transposed view of an object-major buffer."""
    t = torch.from_numpy(np.ascontiguousarray(np.swapaxes(a, 0, 1))).to(device=dev, dtype=dtype)
    return t.transpose(0, 1)


def run_loss(dev, dtype, training=True, init=None, **kw):
    lw = layer_view(G["logw"], dev, dtype).detach().requires_grad_(True)
    ct = torch.from_numpy(G["cost_target"]).to(device=dev, dtype=dtype).requires_grad_(True)
    mod = mcl.MonteCarloPoseLoss(init_norm_factor=float(G["init_norm_factor"]) if init is None else init,
                                 momentum=float(G["momentum"]), **kw).to(dev)
    mod.train(training)
    loss = mod(lw, ct, torch.tensor(float(G["norm_in"]), dtype=dtype, device=dev))
    return mod, loss, lw, ct


def check_against_reference(dev, dtype, tol, nan_column_zero):
    mod, loss, lw, ct = run_loss(dev, dtype)
    (loss * float(G["coef"])).backward()
    assert abs(loss.item() - float(G["ref64_loss"])) < tol * abs(float(G["ref64_loss"]))
    assert abs(mod.norm_factor.item() - float(G["ref64_norm_factor_after"])) < 1e-6
    g = lw.grad.cpu().double().numpy()
    ref = G["ref64_grad_logw"]
    assert np.abs(g[:, FINITE] - ref[:, FINITE]).max() < tol * np.abs(ref[:, FINITE]).max()
    # the masked object: the reference's in-place `loss_pose[isnan] = 0` leaves 0 * NaN = NaN in this column; the native
    # backward writes exact zeros (the mask's evident intent), the composite inherits torch's NaN
    assert np.isnan(ref[:, 5]).all()
    assert (g[:, 5] == 0).all() if nan_column_zero else (np.isnan(g[:, 5]) | (g[:, 5] == 0)).all()
    np.testing.assert_allclose(ct.grad.cpu().double().numpy(), G["ref64_grad_cost_target"], rtol=tol, atol=0)
    after = float(G["ref64_norm_factor_after"])              # the golden's eval pass ran on the updated module
    mod, loss, _, _ = run_loss(dev, dtype, training=False, init=after)
    assert abs(mod.norm_factor.item() - after) < 1e-7        # eval mode: no EMA update
    assert abs(loss.item() - float(G["ref64_loss_eval"])) < tol * abs(float(G["ref64_loss_eval"]))


def test_composite_loss_matches_the_reference_run_fp64():
    check_against_reference(torch.device("cpu"), torch.float64, 1e-12, nan_column_zero=False)


def test_composite_loss_matches_the_reference_run_fp32():
    check_against_reference(torch.device("cpu"), torch.float32, 2e-6, nan_column_zero=False)


def test_native_loss_matches_the_reference_run(dev, native_on):
    check_against_reference(dev, torch.float32, 2e-6, nan_column_zero=True)


def test_native_lse_and_weights_at_the_infinities(dev, native_on):
    lw = layer_view(G["logw_edge"], dev)
    lse = mcl.mc_logsumexp(lw).cpu().double().numpy()
    ref = G["ref64_lse_edge"]
    assert np.isnan(lse[5]) and np.isnan(ref[5])
    assert lse[3] == -np.inf and ref[3] == -np.inf and lse[7] == np.inf and ref[7] == np.inf
    ok = np.isfinite(ref)
    assert np.abs(lse[ok] - ref[ok]).max() < 2e-6 * np.abs(ref[ok]).max()
    w = mcl.mc_sample_weights(lw)
    assert w.shape == lw.shape and w.stride() == lw.stride()
    w = w.cpu().double().numpy()
    rw = G["ref64_weights_edge"]
    assert (np.isnan(w) == np.isnan(rw)).all()                   # NaN object, the all -inf object and the +inf object
    assert np.nanmax(np.abs(w - rw)) < 2e-6


@pytest.mark.parametrize("D", [4, 7])
def test_native_score_te(dev, native_on, D):
    lw = layer_view(G["logw"], dev)
    smp = layer_view(G[f"samples_d{D}"], dev)
    opt = torch.from_numpy(G[f"pose_opt_d{D}"]).to(device=dev, dtype=torch.float32)
    got = mcl.mc_score_te(smp, opt, lw).cpu().double().numpy()
    ref = G[f"restated_score_te_d{D}"]
    assert np.isnan(got[5]) and np.isnan(ref[5])
    assert np.abs(got[FINITE] - ref[FINITE]).max() < 5e-6


@pytest.mark.parametrize("D", [4, 7])
def test_composite_score_te_and_weights(D):
    dev = torch.device("cpu")
    lw = layer_view(G["logw"], dev, torch.float64)
    got = mcl.mc_score_te(layer_view(G[f"samples_d{D}"], dev, torch.float64), torch.from_numpy(G[f"pose_opt_d{D}"]), lw)
    np.testing.assert_allclose(got.numpy()[FINITE], G[f"restated_score_te_d{D}"][FINITE], rtol=1e-12)
    np.testing.assert_allclose(mcl.mc_sample_weights(lw).numpy()[:, FINITE], G["restated_weights"][:, FINITE], rtol=1e-12)


def test_detection_flavour_reductions(dev, native_on):
    """mmdet's weighted_loss semantics around the per-object loss (Det monte_carlo_pose_loss.py:12-66)."""
    B = G["logw"].shape[1]
    weight = torch.linspace(0.5, 1.5, B, device=dev)
    weight[5] = 0.0
    per_obj = G["cost_target"] + G["ref64_lse"]
    per_obj[5] = 0.0
    wn = weight.cpu().double().numpy()
    nf = float(G["init_norm_factor"])
    mod = mcl.MonteCarloPoseLoss(loss_weight=0.5, init_norm_factor=nf, reduction='mean').to(dev).eval()
    lw = layer_view(G["logw"], dev)
    ct = torch.from_numpy(G["cost_target"]).to(device=dev, dtype=torch.float32)
    nrm = torch.tensor(1.0, device=dev)
    f = lambda **kw: mod(lw, ct, nrm, **kw).cpu().double().numpy()
    np.testing.assert_allclose(f(weight=weight), (per_obj * wn).mean() * 0.5 / nf, rtol=3e-6)
    np.testing.assert_allclose(f(weight=weight, avg_factor=7.0), (per_obj * wn).sum() / 7.0 * 0.5 / nf, rtol=3e-6)
    np.testing.assert_allclose(f(reduction_override='sum'), per_obj.sum() * 0.5 / nf, rtol=3e-6)
    np.testing.assert_allclose(f(reduction_override='none'), per_obj * 0.5 / nf, rtol=3e-6, atol=1e-6)
    with pytest.raises(ValueError):
        f(avg_factor=3.0, reduction_override='sum')


def test_epilogue_argument_checks(dev):
    lw = torch.zeros(3, 8, device=dev)
    with pytest.raises(ValueError):
        native.mc_epilogue(lw, want_lse=False, want_score=True)
    with pytest.raises(native.NativeError):
        native.mc_epilogue(lw, want_lse=False)                   # no output requested -> EPNP_ERR_BAD_ARG
    assert native.mc_epilogue(torch.zeros(0, 8, device=dev))["lse"].shape == (0,)
