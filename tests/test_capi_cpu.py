"""The C-ABI shared library loads on a CPU-only box and exports every symbol include/epropnp_b200.h
declares; argument validation that needs no device works; the product path refuses CPU tensors."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT
from epropnp_b200 import build, capi, native


@pytest.fixture(scope="module")
def handle():
    build.build_library()
    return capi.lib()


def test_header_symbols_are_exported(handle):
    header = open(os.path.join(ROOT, "include", "epropnp_b200.h")).read()
    declared = set(re.findall(r"\b(epnp_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    assert declared == set(capi.exported_symbols()), declared ^ set(capi.exported_symbols())
    for name in declared:
        assert hasattr(handle, name), name


def test_abi_basics(handle):
    assert handle.epnp_abi_version() == capi.ABI_VERSION == 2
    assert handle.epnp_error_string(0) == b"ok"
    assert b"shared memory" in handle.epnp_error_string(-2)
    p = capi.default_params(6)
    assert (p.dof, p.lm_iter, p.mc_samples, p.mc_iter, p.acg_mle_iter) == (6, 10, 512, 4, 3)
    assert abs(p.initial_radius - 30.0) < 1e-6 and abs(p.eps - 1e-5) < 1e-12 and abs(p.z_min - 0.1) < 1e-7
    assert ctypes.sizeof(capi.EpnpParams) == 64
    # resident-point capacity: the bench / dense configs must fit
    assert handle.epnp_max_points(6, 512, 4) >= 4096
    assert handle.epnp_max_points(6, 0, 0) >= 4096
    assert handle.epnp_max_points(6, 512, 4) % 4 == 0


def test_bad_arguments_are_rejected_without_a_device(handle):
    p = capi.default_params(6)
    null = None
    assert handle.epnp_lm_solve_f32(null, null, null, null, null, null, null, null, null, null, null, null, null,
                                    4, 64, ctypes.byref(p), None) == -1
    p.mc_samples, p.mc_iter = 510, 4       # M % I != 0
    one = ctypes.c_void_p(16)
    assert handle.epnp_amis_f32(one, one, one, one, null, null, one, one, one, null, null, null, 0, 0, one, one, null,
                                4, 64, ctypes.byref(p), None) == -1
    p4 = capi.default_params(4)
    p4.mc_iter = 9                           # more AMIS iterations than the kernel's proposal table holds
    assert handle.epnp_amis_f32(one, one, one, one, null, null, one, one, one, null, null, null, 0, 0, one, one, null,
                                4, 64, ctypes.byref(p4), None) == -1
    assert handle.epnp_fused_workspace_bytes(4096, 512, ctypes.byref(capi.default_params(6))) > 4096 * 512 * 28


def test_product_path_refuses_cpu_tensors():
    from epropnp.camera import PerspectiveCamera
    from epropnp.cost_fun import HuberPnPCost
    from epropnp.levenberg_marquardt import LMSolver
    x3d, x2d, w2d = torch.rand(2, 8, 3), torch.rand(2, 8, 2), torch.rand(2, 8, 2)
    cam = PerspectiveCamera(cam_mats=torch.eye(3).expand(2, 3, 3))
    pose = torch.tensor([[0, 0, 5, 1, 0, 0, 0.]]).repeat(2, 1)
    with pytest.raises(native.NativeError, match="no CPU"):
        LMSolver(dof=6)(x3d, x2d, w2d, cam, HuberPnPCost(), pose_init=pose)


def test_reference_surface_is_importable():
    import epropnp.camera as c, epropnp.common as m, epropnp.cost_fun as f          # noqa: E401
    import epropnp.distributions as d, epropnp.epropnp as e, epropnp.levenberg_marquardt as l   # noqa: E401
    assert "epro-pnp_b200" in os.path.abspath(e.__file__)
    for mod, names in ((e, "EProPnP6DoF EProPnP4DoF EProPnPBase cholesky_wrapper"),
                       (l, "LMSolver RSLMSolver solve_wrapper"), (c, "PerspectiveCamera"),
                       (f, "HuberPnPCost AdaptiveHuberPnPCost huber_kernel huber_d_kernel"),
                       (m, "evaluate_pnp pnp_normalize pnp_denormalize quaternion_to_rot_mat yaw_to_rot_mat skew"),
                       (d, "AngularCentralGaussian VonMisesUniformMix")):
        for n in names.split():
            assert hasattr(mod, n), n
    layer = e.EProPnP6DoF(mc_samples=512, num_iter=4, solver=l.LMSolver(dof=6, num_iter=10))
    assert len(layer.state_dict()) == 0 and layer.iter_samples == 128
    with pytest.raises(AssertionError):
        e.EProPnP6DoF(mc_samples=510, num_iter=4)
    # empty batch keeps the reference's shapes without touching the device
    cam = c.PerspectiveCamera(cam_mats=torch.zeros(0, 3, 3))
    r = layer.monte_carlo_forward(torch.zeros(0, 8, 3), torch.zeros(0, 8, 2), torch.zeros(0, 8, 2), cam,
                                  f.HuberPnPCost(), pose_init=torch.zeros(0, 7), force_init_solve=False)
    assert r[0].shape == (0, 7) and r[3].shape == (512, 0, 7) and r[4].shape == (512, 0) and r[5].shape == (0,)
    s = l.LMSolver(dof=6).solve(torch.zeros(0, 8, 3), torch.zeros(0, 8, 2), torch.zeros(0, 8, 2), cam,
                                f.HuberPnPCost(), pose_init=torch.zeros(0, 7), with_pose_cov=True, with_cost=True)
    assert s[0].shape == (0, 7) and s[1].shape == (0, 6, 6) and s[2].shape == (0,)


def test_torch_side_helpers_match_oracle():
    from oracle import pnp_oracle as orc
    from epropnp.common import pnp_denormalize, pnp_normalize, quaternion_to_rot_mat, skew, yaw_to_rot_mat
    from epropnp.camera import PerspectiveCamera
    g = torch.Generator().manual_seed(3)
    q = torch.randn(5, 4, generator=g, dtype=torch.float64)
    q = q / q.norm(dim=-1, keepdim=True)
    assert torch.allclose(quaternion_to_rot_mat(q), orc.quat_to_rotmat(q))
    yaw = torch.randn(5, generator=g, dtype=torch.float64)
    assert torch.allclose(yaw_to_rot_mat(yaw), orc.yaw_to_rotmat(yaw))
    v = torch.randn(5, 3, generator=g, dtype=torch.float64)
    w = torch.randn(5, 3, generator=g, dtype=torch.float64)
    assert torch.allclose((skew(v) @ w[..., None]).squeeze(-1), torch.linalg.cross(v, w))
    assert torch.allclose(PerspectiveCamera.get_quaternion_transfrom_mat(q), orc.quat_tangent_map(q))
    x3d = torch.randn(5, 9, 3, generator=g, dtype=torch.float64)
    pose = torch.cat((torch.randn(5, 3, generator=g, dtype=torch.float64), q), -1)
    off, xn, pn = pnp_normalize(x3d, pose)
    off_o, xn_o, pn_o = orc.normalize_points(x3d, pose)
    assert torch.allclose(xn, xn_o) and torch.allclose(pn, pn_o)
    assert torch.allclose(pnp_denormalize(off, pn), pose)
    # standalone projection utility against the oracle's evaluate (camera Jacobian * weights)
    cam_mats = torch.tensor([[800., 0, 320], [0, 800, 240], [0, 0, 1]], dtype=torch.float64).expand(5, 3, 3)
    pose[:, 2] += 6
    cam = PerspectiveCamera(cam_mats=cam_mats, z_min=0.1)
    u, jac = cam.project(x3d, pose, out_jac=True)
    x2d = u + 0.5
    w2d = torch.ones_like(x2d)
    e = orc.evaluate(x3d, x2d, w2d, pose, orc.Camera(cam_mats, 0.1), 1e9, want_jac=True)
    assert torch.allclose(jac.flatten(-3, -2), e["jac"], rtol=1e-9, atol=1e-9)


def test_distribution_and_cost_classes_match_oracle():
    """Public utility classes (CPU-capable torch code) against the pinned oracle formulas."""
    from oracle import pnp_oracle as orc
    from epropnp.cost_fun import AdaptiveHuberPnPCost, HuberPnPCost
    from epropnp.distributions import AngularCentralGaussian, VonMisesUniformMix
    g = torch.Generator().manual_seed(11)
    d = torch.float64
    A = torch.randn(5, 4, 4, generator=g, dtype=d)
    L = torch.linalg.cholesky(A @ A.transpose(-1, -2) + 0.2 * torch.eye(4, dtype=d))
    x = torch.randn(7, 5, 4, generator=g, dtype=d)
    x = x / x.norm(dim=-1, keepdim=True)
    acg = AngularCentralGaussian(L)
    assert torch.allclose(acg.log_prob(x), orc.acg_logpdf(x, L), rtol=1e-10, atol=1e-10)
    s = acg.sample((1000,))
    assert s.shape == (1000, 5, 4) and torch.allclose(s.norm(dim=-1), torch.ones(1000, 5, dtype=d))
    loc, kappa = torch.randn(5, generator=g, dtype=d), torch.rand(5, generator=g, dtype=d) * 30 + 0.1
    yaw = (torch.rand(9, 5, generator=g, dtype=d) * 2 - 1) * 3.14159
    vm = VonMisesUniformMix(loc, kappa)
    assert torch.allclose(vm.log_prob(yaw), orc.vm_mix_logpdf(yaw, loc, kappa), rtol=1e-9, atol=1e-9)
    smp = vm.sample((64,))
    assert smp.shape == (64, 5) and smp.abs().max() <= 3.1416
    # HuberPnPCost.compute against the oracle's evaluate (cost, rescaled residual, rescaled Jacobian)
    B, N = 3, 11
    x3d = torch.randn(B, N, 3, generator=g, dtype=d)
    pose = torch.tensor([[0.1, -0.2, 6.0, 1.0, 0.0, 0.0, 0.0]], dtype=d).repeat(B, 1)
    K = torch.tensor([[800., 0, 320], [0, 800, 240], [0, 0, 1]], dtype=d).expand(B, 3, 3)
    from epropnp.camera import PerspectiveCamera
    cam = PerspectiveCamera(cam_mats=K, z_min=0.1)
    u, jac_cam = cam.project(x3d, pose, out_jac=True)
    x2d = u + 3 * torch.randn(B, N, 2, generator=g, dtype=d)
    w2d = torch.rand(B, N, 2, generator=g, dtype=d) + 0.5
    cf = AdaptiveHuberPnPCost(relative_delta=0.3)
    cf.set_param(x2d, w2d)
    assert torch.allclose(cf.delta, orc.adaptive_delta(x2d, w2d, 0.3))
    res, cost, jac = cf.compute(u, x2d, w2d, jac_cam=jac_cam, out_residual=True, out_cost=True, out_jacobian=True)
    e = orc.evaluate(x3d, x2d, w2d, pose, orc.Camera(K, 0.1), cf.delta, want_jac=True)
    assert torch.allclose(cost, e["cost"]) and torch.allclose(res, e["residual"]) and torch.allclose(jac, e["jac"])
    buf = torch.empty(B, dtype=d)
    assert cf.compute(u, x2d, w2d, out_cost=buf)[1] is buf and torch.allclose(buf, e["cost"])
    c2 = cf.shallow_copy().repeat_(2)
    assert c2.delta.shape == (2 * B,) and isinstance(c2, AdaptiveHuberPnPCost)
    assert HuberPnPCost(delta=2.0).shallow_copy().delta == 2.0


def test_plain_c_host_links_and_runs(handle, tmp_path):
    """The boundary is a C ABI: a C99 program built with gcc against include/epropnp_b200.h links to the library
    and exercises the device-free entry points."""
    import subprocess
    exe = str(tmp_path / "host_c")
    lib_dir = os.path.dirname(capi.lib_path())
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "examples", "host_c.c"), "-o", exe, "-L", lib_dir, "-lepropnp_b200",
                    "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr


def test_validated_kernels_are_bit_identical():
    """The default build's kernels that were validated on hardware must not change without a new GPU validation:
    experiments sit behind build options, new entry points bring their own kernels (tools/sass_identity.py)."""
    import importlib.util
    import shutil
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    spec = importlib.util.spec_from_file_location("sass_identity", os.path.join(ROOT, "tools", "sass_identity.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    changed, _ = mod.compare(build.build_library())
    assert changed == []


def test_problem_rejects_wrong_shapes():
    import torch
    from epropnp_b200 import native

    class _T(torch.Tensor):
        pass
    x3d, x2d, w2d = torch.zeros(2, 8, 3), torch.zeros(2, 8, 2), torch.zeros(2, 8, 2)
    orig = native._need_cuda
    native._need_cuda = lambda t, what: None
    try:
        cam = torch.eye(3).expand(2, 3, 3)
        with pytest.raises(ValueError):
            native.Problem(x3d, torch.zeros(2, 7, 2), w2d, cam, None, None, 1.0)          # N mismatch
        with pytest.raises(ValueError):
            native.Problem(torch.zeros(2, 8, 2), x2d, w2d, cam, None, None, 1.0)          # x3d last dim
        with pytest.raises(ValueError):
            native.Problem(x3d, x2d, torch.zeros(2, 8, 3), cam, None, None, 1.0)          # w2d last dim
        p = native.Problem(x3d, x2d, torch.ones(2, 8, 1), cam, None, None, 1.0)           # broadcast weight is expanded
        assert tuple(p.w2d.shape) == (2, 8, 2) and p.w2d.is_contiguous()
        with pytest.raises(ValueError):
            native.lm_solve(p, torch.zeros(2, 4), native.default_params(6))                # pose of the wrong width
    finally:
        native._need_cuda = orig
