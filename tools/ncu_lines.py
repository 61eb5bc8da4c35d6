#!/usr/bin/env python3
"""Per-source-line view of an ncu capture.

`ncu --page source --csv` prints per-SASS-instruction metrics; `nvdisasm -gi` of the same cubin gives, for every
instruction, the source line and the chain of call sites it was inlined through.  This tool joins the two by instruction
order and aggregates warp instructions, thread instructions and stall samples
  * by the line of the KERNEL BODY an instruction belongs to (outermost frame in the kernel's own file), and
  * by innermost (file, line).
Usage:
  tools/ncu_lines.py <report.ncu-rep> <object.o> <kernel-name-substring> [--top N] [--json out.json]
"""
from __future__ import annotations

import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_metrics(rep: str, kernel: str):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    name = rows[hdr_i - 1][1] if hdr_i else ""
    col = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[hdr_i + 1:]:
        if len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        def num(k):
            try:
                return float(r[col[k]])
            except (KeyError, ValueError):
                return 0.0
        recs.append({"sass": r[col["Source"]].strip(), "samples": num("# Samples"), "inst": num("Instructions Executed"),
                     "tinst": num("Thread Instructions Executed"), "tags": num("L1 Tag Requests Global"),
                     "long_sb": num("stall_long_sb"), "wait": num("stall_wait"), "branch": num("stall_branch_resolving"),
                     "short_sb": num("stall_short_sb"), "math": num("stall_math"), "no_inst": num("stall_no_inst")})
    return name, recs


def line_info(obj: str, kernel: str):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-gi", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l)
    out, chain, pending = [], [], []
    fl = re.compile(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?')
    ins = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(.*?);")
    for l in lines[start + 1:]:
        if l.startswith(".text.") or l.startswith(".section"):
            break
        m = fl.search(l)
        if m:
            if not pending or pending[-1][2] != (m.group(1), int(m.group(2))):
                pending = [(m.group(1), int(m.group(2)), (m.group(3), int(m.group(4))) if m.group(3) else None)]
            else:
                pending.append((m.group(1), int(m.group(2)), (m.group(3), int(m.group(4))) if m.group(3) else None))
            continue
        m = ins.match(l)
        if m:
            if pending:
                chain = [(p[0], p[1]) for p in pending]
                if pending[-1][2]:
                    chain.append(pending[-1][2])
                pending = []
            out.append((m.group(2).strip(), list(chain)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--json")
    a = ap.parse_args()
    name, recs = sass_metrics(a.report, a.kernel)
    info = line_info(a.obj, a.kernel)
    if len(info) != len(recs):
        print(f"warning: {len(recs)} profiled instructions vs {len(info)} disassembled — object and report differ?", file=sys.stderr)
    n = min(len(info), len(recs))
    kfile = None
    for _, ch in info:
        if ch:
            kfile = ch[-1][0]
            break
    tot = defaultdict(float)
    by_outer, by_inner = defaultdict(lambda: defaultdict(float)), defaultdict(lambda: defaultdict(float))
    for i in range(n):
        r, (_, ch) = recs[i], info[i]
        outer = next(((f, ln) for f, ln in reversed(ch) if f == kfile), ch[-1] if ch else ("?", 0))
        # outermost frame in the kernel's file that lies inside the kernel body: take the LAST chain element
        outer = ch[-1] if ch else ("?", 0)
        inner = ch[0] if ch else ("?", 0)
        for k in ("samples", "inst", "tinst", "tags", "long_sb", "wait", "branch", "short_sb", "math", "no_inst"):
            by_outer[outer][k] += r[k]
            by_inner[inner][k] += r[k]
            tot[k] += r[k]
    def table(d, title):
        print(f"\n== {title} ==  (share of warp instructions | lanes | share of stall samples | long_sb wait branch)")
        rows = sorted(d.items(), key=lambda kv: -kv[1]["inst"])[: a.top]
        for (f, ln), v in rows:
            lanes = v["tinst"] / v["inst"] if v["inst"] else 0
            print(f"{os.path.basename(f)}:{ln:<5d} inst {100 * v['inst'] / tot['inst']:5.1f}%  lanes {lanes:4.1f}  samples {100 * v['samples'] / max(tot['samples'], 1):5.1f}%"
                  f"  long_sb {100 * v['long_sb'] / max(tot['samples'], 1):4.1f}% wait {100 * v['wait'] / max(tot['samples'], 1):4.1f}% branch {100 * v['branch'] / max(tot['samples'], 1):4.1f}%"
                  f"  tags {100 * v['tags'] / max(tot['tags'], 1):4.1f}%")
    print(f"kernel: {name}\ninstructions {n}; warp inst {tot['inst']:.4g}; thread inst {tot['tinst']:.4g}; lanes {tot['tinst'] / max(tot['inst'], 1):.2f}; samples {tot['samples']:.0f}")
    table(by_outer, "by kernel-body line")
    table(by_inner, "by innermost line")
    if a.json:
        js = {"kernel": name, "warp_inst": tot["inst"], "thread_inst": tot["tinst"], "lanes": tot["tinst"] / max(tot["inst"], 1),
              "by_kernel_line": [{"line": f"{os.path.basename(f)}:{ln}", "inst_share": v["inst"] / tot["inst"],
                                  "lanes": v["tinst"] / v["inst"] if v["inst"] else 0, "sample_share": v["samples"] / max(tot["samples"], 1)}
                                 for (f, ln), v in sorted(by_outer.items(), key=lambda kv: -kv[1]["inst"])[: a.top]]}
        with open(a.json, "w") as fh:
            json.dump(js, fh, indent=1)


if __name__ == "__main__":
    main()
