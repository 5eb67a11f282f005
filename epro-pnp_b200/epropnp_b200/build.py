"""Builds the in-tree native artefacts.

    libepropnp_b200.so   nvcc, sm_100a only (cross-compiles without a GPU)
    libhost_emul.so      g++ build of csrc/pnp_math.cuh + tests/host_emul.cpp (CPU tests only)

The .so files stay in the tree (epro-pnp_b200/lib/, git-ignored) so they travel to the GPU box.
"""
import os
import shutil
import subprocess

PKG_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))      # .../epro-pnp_b200
REPO_ROOT = os.path.dirname(PKG_ROOT)
CSRC = os.path.join(PKG_ROOT, "csrc")
LIB_DIR = os.path.join(PKG_ROOT, "lib")
INCLUDE = os.path.join(REPO_ROOT, "include")
LIB_PATH = os.path.join(LIB_DIR, "libepropnp_b200.so")
EMUL_PATH = os.path.join(LIB_DIR, "libhost_emul.so")

# Build options of the SHIPPED library (see EXPERIMENTS below).  Empty: the round-1 kernels as validated on hardware.
# Adopting a measured variant = listing its options here and rewriting profiles/validated_sass.json in the same commit
# (tools/first_gpu_calls.sh revalidate); the CPU emulation and the profiling build follow this list.
DEFAULT_OPTIONS = []

NVCC_FLAGS = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
              "-Xcompiler", "-fPIC", "-shared", "-I", INCLUDE] + DEFAULT_OPTIONS


# Build-option experiments of the kernels: off in the shipped build; tools/variants.py builds
# and A/Bs them on a GPU box, tests/test_simt_emul_cpu.py runs them under the CPU SIMT emulator.
EXPERIMENTS = {
    "lm_packed": ["-DEPNP_LM_PACKED"],
    "sweep_rsq": ["-DEPNP_SWEEP_RSQ"],
    "sweep_noclamp": ["-DEPNP_SWEEP_NOCLAMP"],
    "sweep_split": ["-DEPNP_SWEEP_SPLIT"],
    "sweep_split_noclamp": ["-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP"],
    "all": ["-DEPNP_LM_PACKED", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP"],
    "lm_norefine": ["-DEPNP_LM_NOREFINE"],
    "lm_cost_first": ["-DEPNP_LM_COST_FIRST"],
    "fast_blocksum": ["-DEPNP_FAST_BLOCKSUM"],
    "tf32x3_numerics": ["-DEPNP_TF32X3_NUMERICS"],        # accuracy study of the tensor-core plan, not a speed-up
    "amis_lse": ["-DEPNP_AMIS_LSE"],
    "alias_stage": ["-DEPNP_ALIAS_STAGE"],
    # five resident CTAs per SM: 96 registers (the packed LM evaluation would spill) and 39.8 KB of shared memory per CTA
    "five_ctas": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP",
                  "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE", "-DEPNP_CTAS_PER_SM=5"],
    # six resident CTAs: 80 registers and, without the log-weight buffer, 37.0 KB of shared memory per CTA
    "six_ctas": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP",
                 "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE", "-DEPNP_NO_LW", "-DEPNP_CTAS_PER_SM=6"],
    "no_lw": ["-DEPNP_NO_LW"],
    # the projection on the tensor pipe through the legacy mma.sync m16n8k8 TF32 instruction (3xTF32), 4 CTAs/SM
    "sweep_mma": ["-DEPNP_SWEEP_MMA"],
    "sweep_mma_all": ["-DEPNP_SWEEP_MMA", "-DEPNP_SWEEP_NOCLAMP", "-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST",
                      "-DEPNP_FAST_BLOCKSUM"],
    # ... at five CTAs per SM: 96 registers; the staging ring holds the K[R|t] table, so no aliasing -- 44.9 KB per CTA
    "five_ctas_mma": ["-DEPNP_SWEEP_MMA", "-DEPNP_SWEEP_NOCLAMP", "-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST",
                      "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_NO_LW", "-DEPNP_CTAS_PER_SM=5"],
    # ... and at six: the K[R|t] columns are computed per work item (no table, the ring is aliased away), 80 registers
    "six_ctas_mma": ["-DEPNP_SWEEP_MMA", "-DEPNP_SWEEP_NOCLAMP", "-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST",
                     "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE", "-DEPNP_NO_LW", "-DEPNP_CTAS_PER_SM=6"],
    "sweep_huber_m": ["-DEPNP_SWEEP_HUBER_M"],            # shipped sweep arithmetic with the select-free Huber only
    "six_ctas_huber_m": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE",
                         "-DEPNP_ALIAS_STAGE", "-DEPNP_NO_LW", "-DEPNP_SWEEP_HUBER_M", "-DEPNP_CTAS_PER_SM=6"],
    # the same residency with the shipped sweep arithmetic (18 packed FP ops + 4 MUFU per pair-sample instead of 20 + 2):
    # separates "more resident CTAs" from "different sweep formula" in the A/B
    "five_ctas_plain_sweep": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE",
                              "-DEPNP_ALIAS_STAGE", "-DEPNP_CTAS_PER_SM=5"],
    "six_ctas_plain_sweep": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE",
                             "-DEPNP_ALIAS_STAGE", "-DEPNP_NO_LW", "-DEPNP_CTAS_PER_SM=6"],
    "four_ctas_same_code": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP",
                            "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE"],
    "everything": ["-DEPNP_LM_PACKED", "-DEPNP_LM_NOREFINE", "-DEPNP_LM_COST_FIRST", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP",
                   "-DEPNP_FAST_BLOCKSUM"],
    # round-2 call 2: the measured winners without the (measured slower) cost-first LM, with / without the packed LM evaluation
    "six_hm_nocf": ["-DEPNP_LM_NOREFINE", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE", "-DEPNP_NO_LW",
                    "-DEPNP_SWEEP_HUBER_M", "-DEPNP_CTAS_PER_SM=6"],
    "six_hm_nocf_packed": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_PACKED", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE",
                           "-DEPNP_NO_LW", "-DEPNP_SWEEP_HUBER_M", "-DEPNP_CTAS_PER_SM=6"],
    "five_hm_nocf": ["-DEPNP_LM_NOREFINE", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE",
                     "-DEPNP_SWEEP_HUBER_M", "-DEPNP_CTAS_PER_SM=5"],
    "five_hm_nocf_packed": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_PACKED", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_AMIS_LSE", "-DEPNP_ALIAS_STAGE",
                            "-DEPNP_SWEEP_HUBER_M", "-DEPNP_CTAS_PER_SM=5"],
    "four_hm_nocf_packed": ["-DEPNP_LM_NOREFINE", "-DEPNP_LM_PACKED", "-DEPNP_FAST_BLOCKSUM", "-DEPNP_SWEEP_HUBER_M"],
    "all_norefine": ["-DEPNP_LM_PACKED", "-DEPNP_LM_NOREFINE", "-DEPNP_SWEEP_SPLIT", "-DEPNP_SWEEP_NOCLAMP"],
}


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libepropnp_b200.so")


def build_library(force=False, verbose=False):
    srcs = [os.path.join(CSRC, "pnp_kernels.cu")]
    deps = srcs + [os.path.join(CSRC, "pnp_math.cuh"), os.path.join(INCLUDE, "epropnp_b200.h")]
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or _newer(LIB_PATH, deps):
        cmd = [_nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + srcs
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stderr)
    return LIB_PATH


def build_host_emul(force=False):
    src = os.path.join(REPO_ROOT, "tests", "host_emul.cpp")
    deps = [src, os.path.join(CSRC, "pnp_math.cuh"), os.path.join(INCLUDE, "epropnp_b200.h")]
    os.makedirs(LIB_DIR, exist_ok=True)
    if force or _newer(EMUL_PATH, deps):
        cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-I", INCLUDE, "-o", EMUL_PATH, src]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("g++ failed:\n" + r.stdout + r.stderr)
    return EMUL_PATH


if __name__ == "__main__":
    print(build_library(force=True, verbose=True))
    print(build_host_emul(force=True))
