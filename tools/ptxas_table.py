#!/usr/bin/env python
"""Registers / stack / spills per kernel from the `-Xptxas -v` logs the Makefile keeps next to the objects."""
import glob, os, re, subprocess, sys
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "snch-lbvh_b200", "csrc")
pat = sys.argv[1] if len(sys.argv) > 1 else ""
for log in sorted(glob.glob(os.path.join(root, "*.ptxas.log"))):
    t = open(log).read()
    for m in re.finditer(r"Compiling entry function '(\S+)' for 'sm_100a'\n.*?\n\s+(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads\n.*?Used (\d+) registers(.*)", t):
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        if pat in name:
            print(f"{os.path.basename(log)[:-10]:14s} {name:58s} regs {m.group(5):>3s} stack {m.group(2):>4s} spill st/ld {m.group(3)}/{m.group(4)} {m.group(6).strip()[:60]}")
