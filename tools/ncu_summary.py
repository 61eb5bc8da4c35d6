#!/usr/bin/env python
"""Condense one gpurun_out/<tag>/ directory into the tracked evidence under profiles/:

    python tools/ncu_summary.py gpurun_out/r01a profiles/r01a

writes  <out>_launches.csv   per-kernel launch count / total device time / share of the step (from launches.csv)
        <out>_ncu_<kernel>.json   the `ncu --set full` counters the roofline argument uses, one file per .ncu-rep
        <out>_bench.json     the bench line of the same run (a copy)
and refreshes profiles/traffic.json (dram bytes per launch of the dominant kernel; read by bench.py's `roofline.traffic`).
Runs here (no GPU): `ncu -i` only imports the report.
"""
import collections
import csv
import glob
import io
import json
import os
import shutil
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__occupancy_limit_registers", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__sass_average_branch_targets_threads_uniform.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum",
]


def to_us(v, unit):
    v = float(v.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}[unit]


def launches(src, out):
    p = os.path.join(src, "launches.csv")
    if not os.path.exists(p):
        return
    lines = [l for l in open(p) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = row["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += to_us(row["Metric Value"], row["Metric Unit"])
    tot = sum(a[1] for a in agg.values()) or 1.0
    with open(out + "_launches.csv", "w") as f:
        f.write("kernel,launches,total_us,avg_us,share\n")
        for k, a in agg.items():
            f.write(f"\"{k}\",{a[0]},{a[1]:.1f},{a[1] / a[0]:.1f},{a[1] / tot:.5f}\n")
    print(open(out + "_launches.csv").read())


def full(src, out):
    traffic = {}
    for rep in sorted(glob.glob(os.path.join(src, "*.ncu-rep"))):
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        res = []
        for vals in rows[2:]:
            d = {"kernel": vals[hdr.index("Kernel Name")].split("(")[0]}
            for i, h in enumerate(hdr):
                if h in KEYS or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
                    try:
                        d[h] = [float(vals[i].replace(",", "")), units[i]]
                    except ValueError:
                        pass
            res.append(d)
        name = os.path.basename(rep)[:-8]
        json.dump(res, open(f"{out}_ncu_{name}.json", "w"), indent=1)
        d = max(res, key=lambda r: r.get("gpu__time_duration.sum", [0.0, ""])[0] * {"ms": 1e3, "us": 1.0, "ns": 1e-3, "s": 1e6}.get(r.get("gpu__time_duration.sum", [0, "us"])[1], 1.0))  # the full-size launch of the capture
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}
        tb = sum(d[k][0] * scale[d[k][1]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in d)
        import re
        short = re.sub(r"<.*", "", d["kernel"].split("::")[-1]).replace("void ", "").strip()
        traffic[short + "_bytes_per_launch"] = tb
        # per-kernel entry bench.py scales to its own launch: dram bytes, the batch size of the capture (SNCH_NCU_QUERIES, default
        # the 16.7M-query bench batch) and the L1 global-load sectors per query (what the kernel actually pulls through L1)
        nq = int(os.environ.get("SNCH_NCU_QUERIES", str(1 << 24)))
        ent = {"dram_bytes": tb, "queries": nq, "source": os.path.basename(out)}
        sk = "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"
        if sk in d:
            ent["l1_global_load_sectors_per_query"] = d[sk][0] / nq
        traffic[short] = ent
        print(name, {k: v for k, v in d.items() if k in ("kernel", "gpu__time_duration.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                                                           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct")}, "dram bytes", tb)
    if traffic:
        tp = os.path.join(os.path.dirname(out), "traffic.json")
        old = json.load(open(tp)) if os.path.exists(tp) else {}
        old.update(traffic)
        old["source"] = os.path.basename(out)
        json.dump(old, open(tp, "w"), indent=1)


if __name__ == "__main__":
    src, out = sys.argv[1], sys.argv[2]
    launches(src, out)
    full(src, out)
    for f in ("bench.json", "bench_reference.json"):
        if os.path.exists(os.path.join(src, f)) and os.path.getsize(os.path.join(src, f)):
            shutil.copy(os.path.join(src, f), f"{out}_{f}")
