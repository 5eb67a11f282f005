"""The AMIS kernel's push epilogue on ONE GPU: two extra full-batch buffers on the same device stand in for the peers
(the cross-device mapping itself needs two GPUs: tests/test_push_gather_gpu.py).  The rows this "rank" owns must land in
every peer buffer bit for bit, at the global row offset, and nothing else may be touched."""
import pytest
import torch

from epropnp_b200 import native
from epropnp_b200.synth import make_problem

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("dof,N,M", [(6, 64, 128), (4, 51, 96), (6, 2052, 64)])
def test_push_epilogue_with_local_stand_in_peers(cuda_device, dof, N, M):
    dev = cuda_device
    B, off, total = 5, 3, 11
    D = 7 if dof == 6 else 4
    d = {k: v.to(dev) for k, v in make_problem(B, N, seed=60 + dof, dof=dof).items()}
    prob = native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, native.adaptive_delta(d["x2d"], d["w2d"], 0.5))
    p = native.default_params(dof, lm_iter=4, mc_samples=M, mc_iter=2)
    ref = native.lm_amis_fused(prob, d["pose_init"], p, seed=7, obj_offset=off, want_cost=True)
    full_lw = [torch.full((total, M), -7.0, device=dev) for _ in range(3)]
    full_ps = [torch.full((total, D), -7.0, device=dev) for _ in range(3)]
    out = native.lm_amis_fused_push(prob, d["pose_init"], p, full_ps[0][off:off + B], full_lw[0][off:off + B], full_lw[1:], full_ps[1:],
                                    seed=7, obj_offset=off, want_cost=True, want_cov=True)
    torch.cuda.synchronize()
    for k in ("pose_samples", "cost"):
        assert torch.equal(out[k], ref[k]), k
    for r in range(3):
        assert torch.equal(full_lw[r][off:off + B], ref["logw"]) and torch.equal(full_ps[r][off:off + B], ref["pose_opt"])
        for t in (full_lw[r], full_ps[r]):
            assert (t[:off] == -7.0).all() and (t[off + B:] == -7.0).all()
