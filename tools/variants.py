#!/usr/bin/env python
"""Build-option experiments of the sm_100a kernels: build them, read their SASS, and A/B them on a GPU box.

The default build of csrc/pnp_kernels.cu is the validated product.  Candidate kernel changes live behind
preprocessor options (they compile to nothing when off, so the default SASS is bit-identical) until a GPU run has
shown them parity-green and faster:

    EPNP_LM_PACKED       LM normal equations with the Jacobian's u / v rows in the two lanes of fp32x2 registers
    EPNP_SWEEP_HUBER_M   the shipped sweep (reciprocal + square root) with the select-free Huber only: 17 packed FP ops
    EPNP_SWEEP_RSQ       AMIS cost sweep with one MUFU.RSQ per point (no reciprocal), select-free Huber
    EPNP_SWEEP_NOCLAMP   ... + clamp-free loop when the pose keeps the whole object in front of z_min
    EPNP_SWEEP_MMA       ... the 3x4 projection on the tensor pipe (mma.sync m16n8k8 TF32, error-compensated 3xTF32)
    EPNP_SWEEP_SPLIT     ... + two samples per thread over half of the points each
    EPNP_LM_NOREFINE     LM step from the plain fp32 Cholesky solve (no fp64-residual refinement on the serial lane)
    EPNP_LM_COST_FIRST   LM accept / reject from a cost-only pass; normal equations only for accepted steps
    EPNP_FAST_BLOCKSUM   block reductions of the AMIS refit as transposed butterflies (16-31 shuffles instead of 5 per value)
    EPNP_AMIS_LSE        mixture densities as one running log-sum-exp per sample (-6 KB shared memory at M = 512, I = 4)
    EPNP_ALIAS_STAGE     staging ring inside the sample buffer when every CTA solves one object (-7 KB)
    EPNP_CTAS_PER_SM=5   launch bounds for five resident CTAs (96 registers; needs the two options above: 39.8 KB / CTA)
    EPNP_NO_LW           no log-weight buffer in shared memory (recomputed where needed): 37.0 KB / CTA, six CTAs at 80 registers

    python tools/variants.py build            # every variant -> epro-pnp_b200/lib/variants/ (they travel with gpurun)
    python tools/variants.py static           # registers / spills / hot-loop instruction mix per variant (no GPU)
    python tools/variants.py run [names...]   # GPU box: bench.py per variant, then the parity tests of the faster ones -> gpurun_out/variants.jsonl

`run` swaps each variant in as lib/libepropnp_b200.so for the duration of its tests + bench and restores the
default build afterwards, so tests and bench exercise exactly the code path a default build of that variant would.
"""
import argparse
import collections
import json
import os
import re
import shutil
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "epro-pnp_b200"))
from epropnp_b200 import build as B  # noqa: E402

VARIANTS = collections.OrderedDict([("default", [])] + list(B.EXPERIMENTS.items()))
VDIR = os.path.join(B.LIB_DIR, "variants")
SRC = os.path.join(B.CSRC, "pnp_kernels.cu")


def vpath(name):
    return os.path.join(VDIR, f"libepropnp_b200.{name}.so")


def build_variant(name, force=False):
    os.makedirs(VDIR, exist_ok=True)
    out, log = vpath(name), vpath(name) + ".ptxas.log"
    deps = [SRC, os.path.join(B.CSRC, "pnp_math.cuh"), os.path.join(B.INCLUDE, "epropnp_b200.h")]
    if force or B._newer(out, deps):
        cmd = [B._nvcc()] + B.NVCC_FLAGS + VARIANTS[name] + ["-Xptxas", "-v", "-o", out, SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}{r.stderr}")
        with open(log, "w") as f:
            f.write(r.stderr)
    return out


def ptxas_summary(name, pattern="solve_kernelILi6ELb1ELb1"):
    txt = open(vpath(name) + ".ptxas.log").read()
    for m in re.finditer(r"Compiling entry function '(\S+)'.*?\n.*?Function properties.*?\n\s*(.*?)\n.*?Used (\d+) registers", txt, re.S):
        if pattern in m.group(1):
            sp = re.search(r"(\d+) bytes spill stores, (\d+) bytes spill loads", m.group(2))
            return dict(registers=int(m.group(3)), spill_store_bytes=int(sp.group(1)), spill_load_bytes=int(sp.group(2)))
    return {}


def sass_loops(so, pattern="solve_kernelILi6ELb1ELb1"):
    """Backward-branch loops of the fused 6DoF kernel that hold real FP work: (n_instr, opcode histogram)."""
    txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
    out = []
    for part in re.split(r"\n\s*Function : ", txt)[1:]:
        if pattern not in part.split("\n")[0]:
            continue
        lines = [l for l in part.split("\n") if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l)]

        def op(l):
            t = re.sub(r"/\*[0-9a-fx ]+\*/", "", l).strip().rstrip(";").split()
            return (t[1] if t[0].startswith("@") else t[0]).split(".")[0]

        addr = lambda l: int(re.match(r"\s+/\*([0-9a-f]+)\*/", l).group(1), 16)
        amap = {addr(l): i for i, l in enumerate(lines)}
        for i, l in enumerate(lines):
            m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)\s*;", l)
            if not m:
                continue
            t = int(m.group(1), 16)
            if t < addr(l) and t in amap:
                body = lines[amap[t]:i + 1]
                c = collections.Counter(op(x) for x in body)
                fp = sum(c.get(k, 0) for k in ("FFMA", "FFMA2", "FMUL", "FMUL2", "FADD", "FADD2"))
                if c.get("HMMA", 0) and c.get("BAR", 0) == 0 and len(body) < 200:
                    out.append((len(body), dict(c.most_common(8))))
                elif len(body) >= 60 and fp > 30 and c.get("BAR", 0) == 0 and len(body) < 600:
                    out.append((len(body), dict(c.most_common(8))))
        out.insert(0, ("total_sass_instructions", len(lines)))
    return out


def cmd_build(args):
    # nvcc runs are independent: build them side by side (about 30 s each)
    from concurrent.futures import ThreadPoolExecutor
    names = list(args.names or VARIANTS)
    with ThreadPoolExecutor(max_workers=max(1, min(len(names), (os.cpu_count() or 2)))) as pool:
        outs = list(pool.map(lambda n: build_variant(n, force=args.force), names))
    for name, out in zip(names, outs):
        print(out, ptxas_summary(name))


def cmd_static(args):
    for name in (args.names or VARIANTS):
        build_variant(name)
        print(f"== {name}  {' '.join(VARIANTS[name]) or '(default build)'}")
        print("   fused 6DoF kernel:", ptxas_summary(name))
        loops = sass_loops(vpath(name))
        print("  ", loops[0])
        for n, mix in loops[1:]:
            kind = "mma-tile" if mix.get("HMMA", 0) else "sweep" if mix.get("FFMA2", 0) >= 30 and mix.get("MUFU", 0) >= 4 and mix.get("LDS", 0) >= 4 and "FFMA" not in mix \
                else ("lm-eval" if (mix.get("FFMA", 0) + mix.get("FFMA2", 0)) >= 90 else "other")
            if kind != "other":
                print(f"   {kind:8s} loop {n:4d} instr  {mix}")


def cmd_run(args):
    # default selection: every option on its own + everything together (the unions in between are left to the caller)
    # default selection: the candidate combinations first (if the call runs out of time the single options are the ones
    # lost), then every option on its own for attribution
    names = args.names or ["default", "six_ctas_huber_m", "six_ctas_plain_sweep", "six_ctas", "five_ctas_mma", "six_ctas_mma", "sweep_mma_all",
                           "five_ctas_plain_sweep", "five_ctas",
                           "everything", "four_ctas_same_code", "lm_cost_first", "lm_norefine", "sweep_huber_m", "sweep_rsq",
                           "sweep_split", "sweep_noclamp", "fast_blocksum", "lm_packed"]
    os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
    out_path = os.path.join(REPO, "gpurun_out", "variants.jsonl")
    B.build_library()
    # anything missing or stale (e.g. the snapshot did not keep mtimes) is rebuilt side by side, not one per variant later
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=max(1, min(len(names), (os.cpu_count() or 2)))) as pool:
        list(pool.map(build_variant, names))
    keep = B.LIB_PATH + ".default_build"
    shutil.copy(B.LIB_PATH, keep)
    env = dict(os.environ, PYTHONPATH=os.path.join(REPO, "epro-pnp_b200") + os.pathsep + os.environ.get("PYTHONPATH", ""))

    def bench(name):
        vals = []
        for _ in range(args.repeats):
            b = subprocess.run(["timeout", "150", sys.executable, "bench.py", "--steps", str(args.steps), "--warmup",
                                str(args.warmup), "--no-cpu-baseline", "--no-e2e"], cwd=REPO, env=env, capture_output=True, text=True)
            line = [l for l in b.stdout.splitlines() if l.startswith("{")]
            if b.returncode == 0 and line:
                j = json.loads(line[-1])
                vals.append(dict(value=j["value"], ms_per_step=j["ms_per_step"], sm_mhz=j.get("clocks", {}).get("sm_mhz")))
            else:
                vals.append(dict(error=(b.stderr or b.stdout)[-1500:]))
        return vals

    def tests(name, rec):
        t = subprocess.run(["timeout", str(args.test_timeout), sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu"]
                           + args.tests.split(), cwd=REPO, env=env, capture_output=True, text=True)
        rec["tests_rc"] = t.returncode
        rec["tests_tail"] = t.stdout.strip().splitlines()[-1:] if t.stdout.strip() else []
        if t.returncode != 0:
            rec["tests_fail"] = t.stdout[-3000:]

    def emit(rec):
        with open(out_path, "a") as f:
            f.write(json.dumps(rec) + "\n")
        print(json.dumps(rec)[:400], flush=True)

    def best(vals):
        return max((v.get("value", 0.0) for v in vals), default=0.0)

    try:
        # pass 1: one bench line per variant (about 20 s each) -- a variant that is not faster needs no parity run
        speed = {}
        for name in names:
            shutil.copy(build_variant(name), B.LIB_PATH)
            speed[name] = bench(name)
            emit(dict(variant=name, flags=VARIANTS[name], phase="bench", bench=speed[name]))
        base = best(speed.get("default", [])) or min((best(v) for v in speed.values() if best(v) > 0), default=0.0)
        # pass 2: the GPU parity + edge tests, fastest first, for the variants that beat the default build
        winners = sorted((n for n in names if n != "default" and best(speed[n]) > base * args.min_gain),
                         key=lambda n: -best(speed[n]))
        if args.all_tests:
            winners += [n for n in names if n not in winners and n != "default"]
        for name in winners[:args.max_tested] if args.max_tested > 0 else winners:
            rec = dict(variant=name, flags=VARIANTS[name], phase="tests", speedup=best(speed[name]) / base if base else None)
            shutil.copy(build_variant(name), B.LIB_PATH)
            tests(name, rec)
            emit(rec)
        emit(dict(phase="summary", default=base, ranking=[(n, round(best(speed[n]) / base, 4) if base else None)
                                                         for n in sorted(speed, key=lambda n: -best(speed[n]))]))
    finally:
        shutil.copy(keep, B.LIB_PATH)
        os.remove(keep)


if __name__ == "__main__":
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    for name, fn in (("build", cmd_build), ("static", cmd_static), ("run", cmd_run)):
        p = sub.add_parser(name)
        p.add_argument("names", nargs="*", help="subset of: " + " ".join(VARIANTS))
        p.set_defaults(fn=fn)
        if name == "build":
            p.add_argument("--force", action="store_true")
        if name == "run":
            p.add_argument("--tests", default="tests/test_gpu_parity.py tests/test_gpu_edges.py",
                           help="pytest selection run per variant (add '-k golden' for a quicker pass)")
            p.add_argument("--test-timeout", type=int, default=400)
            p.add_argument("--steps", type=int, default=300)
            p.add_argument("--warmup", type=int, default=5)
            p.add_argument("--repeats", type=int, default=1)
            p.add_argument("--min-gain", type=float, default=1.01, help="parity-test only variants at least this much faster than default")
            p.add_argument("--max-tested", type=int, default=8, help="at most this many variants get the parity run (0 = no limit)")
            p.add_argument("--all-tests", action="store_true", help="parity-test the slower variants too (after the faster ones)")
    a = ap.parse_args()
    for n in a.names:
        if n not in VARIANTS:
            ap.error(f"unknown variant {n!r}")
    a.fn(a)
