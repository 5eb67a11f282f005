#!/usr/bin/env python
"""Per-phase cycle breakdown of the AMIS kernel (profiling build with clock64() timers on the serial thread of every
CTA).  Usage on a GPU box:  python tools/phase_profile.py [B] [N] [M]"""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "epro-pnp_b200"))
import torch  # noqa: E402
from epropnp_b200 import build, capi, native  # noqa: E402
from epropnp_b200.synth import make_problem  # noqa: E402

PHASES = ["load + first fit", "draw+sweep", "logp_old", "weights", "refit_sums", "refit_finish", "output"]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    extra = []
    prof_lib = os.path.join(build.LIB_DIR, "libepropnp_b200_prof.so")
    cmd = [build._nvcc()] + build.NVCC_FLAGS + extra + ["-DEPNP_PHASE_TIMERS", "-o", prof_lib, os.path.join(build.CSRC, "pnp_kernels.cu")]
    subprocess.run(cmd, check=True)
    capi._LIB_PATH = prof_lib
    capi._lib = None
    lib = capi.lib()
    lib.epnp_debug_set_phase_buffer.argtypes = [ctypes.c_void_p]
    dev = torch.device("cuda:0")
    pc = make_problem(B, N, seed=7)
    d = {k: v.to(dev) for k, v in pc.items()}
    delta = native.adaptive_delta(d["x2d"], d["w2d"], 0.5)
    prob = native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, delta)
    p = native.default_params(6, mc_samples=M, mc_iter=4)
    native.lm_amis_fused(prob, d["pose_init"], p, seed=1)
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    lib.epnp_debug_set_phase_buffer(ctypes.c_void_p(buf.data_ptr()))
    native.lm_amis_fused(prob, d["pose_init"], p, seed=1)
    torch.cuda.synchronize()
    lib.epnp_debug_set_phase_buffer(None)
    c = buf.cpu().tolist()[:len(PHASES)]
    tot = sum(c)
    print(f"B={B} N={N} M={M}: cycles of the AMIS kernel's serial thread per object, by phase (sum over CTAs / B)")
    for name, v in zip(PHASES, c):
        print(f"  {name:14s} {v / B:10.0f} cycles  {100.0 * v / tot:5.1f} %")
    print(f"  {'total':14s} {tot / B:10.0f} cycles")


if __name__ == "__main__":
    main()
