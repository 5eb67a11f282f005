#!/usr/bin/env python
"""Turn an `ncu --set full` report into the small JSON summaries kept under profiles/ (runs in the build container:
ncu reads reports without a GPU).

    python tools/ncu_summary.py gpurun_out/fused_full.ncu-rep --tag r2 [--kernel solve_kernel] [--write-traffic]

writes profiles/<tag>_fused_ncu_raw_metrics.json = {metric: [value, unit]} for the metrics the design discussion uses
(duration, DRAM bytes, instruction count, issue / pipe utilisation, occupancy limits, bank conflicts) and, with
--write-traffic, refreshes profiles/traffic.json (read by bench.py for `roofline.traffic`).  `ncu -i REP --page raw
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
STALLS = "smsp__average_warp_latency_issue_stalled_"      # ..._<reason>.ratio (warp-state section of --set full)
UNIT_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def raw_rows(rep):
    r = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.exit("ncu failed: " + (r.stderr or r.stdout)[-800:])
    text = r.stdout[r.stdout.index('"ID"'):] if '"ID"' in r.stdout else r.stdout
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
    ap.add_argument("--kernel", default="solve_kernel", help="substring of the kernel name to summarise (first match)")
    ap.add_argument("--write-traffic", action="store_true")
    a = ap.parse_args()
    kernels = per_kernel(raw_rows(a.report))
    hit = [(n, m) for n, m in kernels if a.kernel in n]
    if not hit:
        sys.exit(f"no kernel matching {a.kernel!r}; report holds: {sorted({n for n, _ in kernels})}")
    name, metrics = hit[0]
    keep = {k: metrics[k] for k in KEEP if k in metrics}
    keep.update({k: v for k, v in metrics.items() if k.startswith(STALLS)})
    missing = [k for k in KEEP if k not in metrics]
    out = os.path.join(ROOT, "profiles", f"{a.tag}_fused_ncu_raw_metrics.json")
    with open(out, "w") as f:
        json.dump(dict(kernel=name, **keep), f, indent=1)
    print("wrote", out, f"({len(keep)} metrics; not in the report: {missing})")
    if a.write_traffic:
        traffic = to_bytes(metrics["dram__bytes_read.sum"]) + to_bytes(metrics["dram__bytes_write.sum"])
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump({"fused_dram_bytes_per_launch": traffic,
                       "source": f"profiles/{a.tag}_fused_ncu_raw_metrics.json (ncu --set full, one launch of {name[:60]})"}, f)
        print("traffic.json:", traffic, "bytes per launch")


if __name__ == "__main__":
    main()
