#!/usr/bin/env python
"""Summarise `nvcc -Xptxas -v` output: registers, spills, static smem per kernel."""
import re, subprocess, sys
txt = "".join(open(f).read() for f in sys.argv[1:])
names = re.findall(r"Compiling entry function '(\S+)'", txt)
dem = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.split("\n")
blocks = re.split(r"ptxas info\s+: Compiling entry function", txt)[1:]
for name, blk in zip(dem, blocks):
    sp = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", blk)
    rg = re.search(r"Used (\d+) registers", blk)
    name = re.sub(r"\(rc::FftPass.*", "", name.replace("void rc::", "").replace("rc::", ""))
    print(f"{name:70s} regs={rg.group(1) if rg else '?':>4s} stack={sp.group(1)} spill={sp.group(2)}/{sp.group(3)}")
