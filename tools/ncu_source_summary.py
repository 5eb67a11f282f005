#!/usr/bin/env python
"""Where a kernel's warp-time goes, from the source page of an `ncu --set full --import-source on` report (runs in the
build container).  Finds the SASS loops (backward branches), attributes sampled warp states and executed instructions to
the hottest inner loop and to everything else, and lists the top stall sites outside it.

    python tools/ncu_source_summary.py gpurun_out/r2_amis_full.ncu-rep --out profiles/r2_amis_source_summary.json
"""
import argparse
import collections
import csv
import io
import json
import re
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--out", required=True)
    a = ap.parse_args()
    r = subprocess.run(["ncu", "-i", a.report, "--page", "source", "--csv"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit(r.stderr[-500:])
    rows = list(csv.reader(io.StringIO(r.stdout)))
    kernel = rows[0][1] if rows and len(rows[0]) > 1 else "?"
    hdr, data = rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    addr = [int(x[ix["Address"]], 16) for x in data]
    src = [x[ix["Source"]] for x in data]
    samp = [int(x[ix["# Samples"]] or 0) for x in data]
    inst = [int(x[ix["Instructions Executed"]] or 0) for x in data]
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    amap = {v: i for i, v in enumerate(addr)}
    loops = []
    for i, s in enumerate(src):
        m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16)
            if t < addr[i] and t in amap:
                loops.append((amap[t], i))
    tot_s, tot_i = sum(samp), sum(inst)

    def opmix(lo, hi):
        c = collections.Counter(re.sub(r"^@!?\S+\s+", "", s).split()[0].split(".")[0] for s in src[lo:hi + 1])
        return dict(c.most_common(8))

    def stall_share(sel):
        c = collections.Counter()
        for i, x in enumerate(data):
            if sel(i):
                for h in stalls:
                    c[h.replace("stall_", "")] += int(x[ix[h]] or 0)
        t = sum(c.values()) or 1
        return {k: round(100.0 * v / t, 1) for k, v in c.most_common(9)}

    # the hottest INNER loop: most instructions among loops that contain no other loop
    inner = [(lo, hi) for lo, hi in loops if not any(lo <= l2 and h2 <= hi and (l2, h2) != (lo, hi) for l2, h2 in loops)]
    inner.sort(key=lambda p: -sum(inst[p[0]:p[1] + 1]))
    lo, hi = inner[0]
    out = dict(kernel=kernel, samples=tot_s, warp_instructions=tot_i,
               hottest_inner_loop=dict(sass_instructions=hi - lo + 1, op_mix=opmix(lo, hi),
                                       pct_of_instructions=round(100.0 * sum(inst[lo:hi + 1]) / tot_i, 1),
                                       pct_of_samples=round(100.0 * sum(samp[lo:hi + 1]) / tot_s, 1),
                                       stall_pct=stall_share(lambda i: lo <= i <= hi)),
               everything_else=dict(pct_of_instructions=round(100.0 * (tot_i - sum(inst[lo:hi + 1])) / tot_i, 1),
                                    pct_of_samples=round(100.0 * (tot_s - sum(samp[lo:hi + 1])) / tot_s, 1),
                                    stall_pct=stall_share(lambda i: not (lo <= i <= hi))),
               other_loops=[dict(sass_instructions=h2 - l2 + 1, pct_of_instructions=round(100.0 * sum(inst[l2:h2 + 1]) / tot_i, 1),
                                 pct_of_samples=round(100.0 * sum(samp[l2:h2 + 1]) / tot_s, 1), op_mix=opmix(l2, h2))
                            for l2, h2 in inner[1:8]],
               top_stall_sites_outside_the_loop=[])
    sites = sorted(((samp[i], i) for i in range(len(data)) if not (lo <= i <= hi)), reverse=True)[:10]
    for s, i in sites:
        why = {h.replace("stall_", ""): int(data[i][ix[h]] or 0) for h in stalls if int(data[i][ix[h]] or 0) > 0.25 * max(s, 1)}
        out["top_stall_sites_outside_the_loop"].append(dict(samples=s, pct_of_samples=round(100.0 * s / tot_s, 2), sass=src[i].strip()[:60], stalls=why))
    with open(a.out, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", a.out, "| inner loop:", out["hottest_inner_loop"]["pct_of_instructions"], "% of instructions,",
          out["hottest_inner_loop"]["pct_of_samples"], "% of samples")


if __name__ == "__main__":
    main()
