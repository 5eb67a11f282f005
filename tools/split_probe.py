#!/usr/bin/env python
"""How long do the two halves of the fused solve take on their own?  Times, at the bench shape, the fused LM + AMIS
launch against the LM-only launch (with covariance) followed by the AMIS-only launch, for the library given by
EPNP_LIB (default: the shipped build).   python tools/split_probe.py [B] [N] [M]"""
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "epro-pnp_b200"))
import torch  # noqa: E402
from epropnp_b200 import capi, native  # noqa: E402
from epropnp_b200.synth import make_problem  # noqa: E402


def timed(fn, iters=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in ev:
        a.record(); fn(); b.record()
    torch.cuda.synchronize()
    return statistics.median(a.elapsed_time(b) for a, b in ev)


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 512
    M = int(sys.argv[3]) if len(sys.argv) > 3 else 512
    if os.environ.get("EPNP_LIB"):
        capi._LIB_PATH = os.environ["EPNP_LIB"]
        capi._lib = None
    dev = torch.device("cuda:0")
    sets = []
    for s in range(3):          # rotate input sets (3 x 59 MB > L2 with the outputs)
        d = {k: v.to(dev) for k, v in make_problem(B, N, seed=7 + s).items()}
        delta = native.adaptive_delta(d["x2d"], d["w2d"], 0.5)
        sets.append((native.Problem(d["x3d"], d["x2d"], d["w2d"], d["cam_mats"], None, None, delta), d["pose_init"]))
    p = native.default_params(6, mc_samples=M, mc_iter=4)
    lm0 = native.lm_solve(sets[0][0], sets[0][1], p, want_cov=True)
    k = [0]

    def nxt():
        k[0] = (k[0] + 1) % len(sets)
        return sets[k[0]]

    def fused():
        prob, pi = nxt()
        native.lm_amis_fused(prob, pi, p, seed=1)

    def lm_only():
        prob, pi = nxt()
        native.lm_solve(prob, pi, p, want_cov=True)

    def amis_only():
        prob, pi = nxt()
        native.amis(prob, lm0["pose_opt"], lm0["pose_cov"], p, seed=1)

    def both():
        prob, pi = nxt()
        r = native.lm_solve(prob, pi, p, want_cov=True)
        native.amis(prob, r["pose_opt"], r["pose_cov"], p, seed=1)

    out = dict(lib=os.path.basename(capi._LIB_PATH) if hasattr(capi, "_LIB_PATH") else "shipped", B=B, N=N, M=M,
               fused_ms=timed(fused), lm_only_ms=timed(lm_only), amis_only_ms=timed(amis_only), lm_then_amis_ms=timed(both))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
