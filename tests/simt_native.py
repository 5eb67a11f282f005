"""TEST-ONLY: run epropnp_b200.native (and everything above it) on CPU tensors through a g++ build of the REAL
kernel source executing under the SIMT emulator in tests/simt_emul/ (see the header comment of
tests/simt_emul/cuda_runtime.h for what that does and does not prove).

`install(monkeypatch, flags=())` builds (once per flag set) epro-pnp_b200/csrc/pnp_kernels.cu with
`g++ -DEPNP_SIMT_EMUL <flags>` and points the ctypes layer at it for the duration of one test.  The product never
imports this file; outside these tests CPU tensors are refused and a missing nvcc-built library is an error.
"""
import contextlib
import ctypes
import hashlib
import os
import subprocess

import torch

from epropnp_b200 import build, capi, native

_HERE = os.path.dirname(os.path.abspath(__file__))
_SHIM = os.path.join(_HERE, "simt_emul")
_handles = {}


def build_emulated(flags=()):
    flags = tuple(build.DEFAULT_OPTIONS) + tuple(f for f in flags if f not in build.DEFAULT_OPTIONS)
    tag = hashlib.sha1(" ".join(flags).encode()).hexdigest()[:8] if flags else "default"
    out = os.path.join(build.LIB_DIR, f"libepropnp_simt_{tag}.so")
    src = os.path.join(build.CSRC, "pnp_kernels.cu")
    deps = build.kernel_sources() + [os.path.join(_SHIM, "cuda_runtime.h"), os.path.join(_SHIM, "math_constants.h")]
    os.makedirs(build.LIB_DIR, exist_ok=True)
    if build._newer(out, deps):
        cmd = ["g++", "-std=c++20", "-O2", "-mfma", "-ffp-contract=fast", "-x", "c++", "-DEPNP_SIMT_EMUL", *flags, "-fPIC", "-shared",
               "-I", _SHIM, "-I", build.INCLUDE, "-I", build.CSRC, "-o", out, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ (SIMT emulation build) failed:\n" + r.stdout + r.stderr)
    return out


def handle(flags=()):
    flags = tuple(flags)
    if flags not in _handles:
        h = ctypes.CDLL(build_emulated(flags))
        for name, (res, args) in capi._SIGNATURES.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _handles[flags] = h
    return _handles[flags]


def install(monkeypatch, flags=()):
    monkeypatch.setattr(capi, "_lib", handle(flags))
    monkeypatch.setattr(native, "_need_cuda", lambda t, what: None)
    monkeypatch.setattr(native, "stream_ptr", lambda device=None: None)
    monkeypatch.setattr(torch.cuda, "device", lambda device=None: contextlib.nullcontext())
    monkeypatch.setattr(torch.cuda, "synchronize", lambda device=None: None)
    return torch.device("cpu")
