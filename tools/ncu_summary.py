#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summaries kept under profiles/ (runs in the build container:
ncu reads reports without a GPU).

    python tools/ncu_summary.py gpurun_out/r2_amis_full.ncu-rep --tag r2 --kernel amis_kernel --stats-key amis_kernel@fused --objects 4096

writes profiles/<tag>_<kernel>_ncu_raw_metrics.json = {metric: [value, unit]} for the metrics the design discussion uses
(duration, DRAM bytes, instruction count, issue / pipe utilisation, occupancy limits, bank conflicts, stall reasons) and,
with --stats-key, the kernel's record in profiles/kernel_stats.json (read by bench.py for `roofline.traffic` and the
`issue` block): per-launch warp-instructions and DRAM bytes TOGETHER WITH the SASS fingerprint of that kernel in the
in-tree library -- run it right after the capture, before the library is rebuilt; bench.py refuses the record once the
kernel's SASS has changed.  `ncu -i REP --page raw
--csv` prints either one row per kernel with one column per metric (plus a units row) or one row per (kernel, metric);
both layouts are understood.
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__occupancy_limit_warps",
    "gpc__cycles_elapsed.avg.per_second", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "lts__t_sector_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"            # ..._<reason>_per_issue_active.ratio (warp-state section of --set full)
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def raw_rows(rep):
    """`rep`: an .ncu-rep, or the saved output of `ncu -i REP --page raw --csv` (tools/revalidate.sh keeps only that for the
    kernels off the hot path)."""
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
        if r.returncode != 0:
            sys.exit("ncu failed: " + (r.stderr or r.stdout)[-800:])
        out = r.stdout
    text = out[out.index('"ID"'):] if '"ID"' in out else out
    return list(csv.reader(io.StringIO(text)))


def per_kernel(rows):
    """-> list of (kernel name, {metric: [value, unit]}) in launch order."""
    head = rows[0]
    if "Metric Name" in head and "Metric Value" in head:                      # one row per (kernel, metric)
        i_id, i_k = head.index("ID"), head.index("Kernel Name")
        i_m, i_u, i_v = head.index("Metric Name"), head.index("Metric Unit"), head.index("Metric Value")
        out, order = {}, []
        for row in rows[1:]:
            if len(row) <= i_v:
                continue
            key = row[i_id]
            if key not in out:
                out[key] = (row[i_k], {})
                order.append(key)
            out[key][1][row[i_m]] = [row[i_v], row[i_u]]
        return [out[k] for k in order]
    i_k = head.index("Kernel Name")                                           # one row per kernel, units in row 2
    units = rows[1] if len(rows) > 1 and not rows[1][head.index("ID")].strip().isdigit() else [""] * len(head)
    body = rows[2:] if units is rows[1] else rows[1:]
    res = []
    for row in body:
        if len(row) != len(head):
            continue
        res.append((row[i_k], {head[j]: [row[j], units[j]] for j in range(len(head)) if "__" in head[j]}))
    return res


def to_bytes(value_unit):
    v, u = value_unit
    return float(v.replace(",", "")) * UNIT_BYTES.get(u, 1.0)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("report")
    ap.add_argument("--tag", required=True, help="file prefix under profiles/, e.g. r2")
    ap.add_argument("--kernel", default="amis_kernel", help="substring of the kernel name to summarise (first match)")
    ap.add_argument("--stats-key", default=None, help="record name in profiles/kernel_stats.json, '<kernel>@<bench config>'")
    ap.add_argument("--objects", type=int, default=4096, help="objects the profiled launch processed")
    ap.add_argument("--sass-symbol", default=None, help="substring of the mangled kernel name (default: derived from --kernel: the 6DoF kernels of the N = 512 bench)")
    a = ap.parse_args()
    kernels = per_kernel(raw_rows(a.report))
    hit = [(n, m) for n, m in kernels if a.kernel in n]
    if not hit:
        sys.exit(f"no kernel matching {a.kernel!r}; report holds: {sorted({n for n, _ in kernels})}")
    name, metrics = hit[0]
    keep = {k: metrics[k] for k in KEEP if k in metrics}
    keep.update({k: v for k, v in metrics.items() if k.startswith(STALLS)})
    missing = [k for k in KEEP if k not in metrics]
    out = os.path.join(ROOT, "profiles", f"{a.tag}_{a.kernel}_ncu_raw_metrics.json")
    rel = os.path.relpath(out, ROOT)
    with open(out, "w") as f:
        json.dump(dict(kernel=name, **keep), f, indent=1)
    print("wrote", out, f"({len(keep)} metrics; not in the report: {missing})")
    if a.stats_key:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import sass_identity
        symbol = a.sass_symbol or {"amis_kernel": "amis_kernelILi6ELi128EE", "lm_warp_kernel": "lm_warp_kernelILi6ELb1ELi1EE"}.get(a.kernel, a.kernel)
        fp = {k: v for k, v in sass_identity.fingerprints(sass_identity.DEFAULT_LIB).items() if symbol in k}
        if len(fp) != 1:
            sys.exit(f"--sass-symbol {symbol!r} matches {sorted(fp)}")
        num = lambda k: float(metrics[k][0].replace(",", ""))
        dur = num("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6, "s": 1e3, "second": 1e3}.get(metrics["gpu__time_duration.sum"][1], 1.0)
        rec = dict(kernel=name, sass_symbol=symbol, sass_sha1=list(fp.values())[0]["sha1"], objects_per_launch=a.objects,
                   warp_instr_per_launch=num("smsp__inst_executed.sum"),
                   dram_bytes_per_launch=to_bytes(metrics["dram__bytes_read.sum"]) + to_bytes(metrics["dram__bytes_write.sum"]),
                   duration_ms=dur, issue_active_pct=num("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                   fma_pipe_pct=num("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"),
                   xu_pipe_pct=num("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
                   source=f"{rel} (ncu --set full --clock-control none, one launch)")
        path = os.path.join(ROOT, "profiles", "kernel_stats.json")
        allrec = json.load(open(path)) if os.path.exists(path) else {}
        for key in a.stats_key.split(","):                  # the same kernel serves several bench configs (per-object scaling)
            allrec[key] = rec
        with open(path, "w") as f:
            json.dump(allrec, f, indent=1, sort_keys=True)
        print("kernel_stats.json:", a.stats_key, {k: rec[k] for k in ("warp_instr_per_launch", "dram_bytes_per_launch", "duration_ms")})


if __name__ == "__main__":
    main()
