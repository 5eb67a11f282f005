#!/usr/bin/env python
"""TEST / BASELINE INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference package for the CPU arm of bench.py:

    python oracle/stage_ref.py          # /root/reference/epropnp/*.py  ->  oracle/_ref/epropnp/  (byte copies)

`oracle/_ref/` is git-ignored (the reference's sources never enter this repository's history) but NOT gpurun-ignored:
it travels to the GPU box with the snapshot, where /root/reference does not exist, so `bench.py --impl reference`
and the `cpu_baseline` leg can time the reference's own PyTorch layer (`EProPnP6DoF.monte_carlo_forward`,
epropnp/epropnp.py:87-196) on the box's host cores.  The reference's one un-vendored dependency (pyro-ppl) is served by
oracle/pyro_shim (69 lines, committed; see its docstring).  __graft_entry__.build() calls this when /root/reference is
present; nothing in the product imports oracle/_ref.
"""
import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/epropnp"
DST = os.path.join(HERE, "_ref", "epropnp")


def stage(verbose=True):
    if not os.path.isdir(SRC):
        return None
    os.makedirs(DST, exist_ok=True)
    names = sorted(n for n in os.listdir(SRC) if n.endswith(".py"))
    for n in names:
        s, d = os.path.join(SRC, n), os.path.join(DST, n)
        if not (os.path.exists(d) and filecmp.cmp(s, d, shallow=False)):
            shutil.copyfile(s, d)
    with open(os.path.join(HERE, "_ref", "STAGED_FROM"), "w") as f:
        f.write(SRC + "\n" + "\n".join(names) + "\n")
    if verbose:
        print(f"staged {len(names)} files of the unmodified reference into {DST}")
    return DST


def staged_path():
    """oracle/_ref (the directory to put on sys.path) if the reference has been staged, else None."""
    return os.path.join(HERE, "_ref") if os.path.exists(os.path.join(DST, "epropnp.py")) else None


if __name__ == "__main__":
    sys.exit(0 if stage() else 1)
