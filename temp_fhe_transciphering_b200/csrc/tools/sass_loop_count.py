#!/usr/bin/env python
"""Development tool: static estimate of the instructions a warp executes per blind-rotation step, from the SASS of a kernel.
Finds the two outermost nested backward branches that enclose the FP64 work (outer = step loop, inner = the rolled
polynomial loop, weight 3), ignores small spin loops (mbarrier waits) and the out-of-line shuffle fallbacks, and prints
the weighted opcode histogram.  usage: sass_loop_count.py <object or .so> <mangled kernel name> [inner weight]"""
import collections
import re
import subprocess
import sys

obj, fun = sys.argv[1], sys.argv[2]
w_inner = int(sys.argv[3]) if len(sys.argv) > 3 else 3
out = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for l in out.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(3), m.group(4)))
addr = [a for a, _, _ in ins]
loops = []
for a, op, rest in ins:
    if op.startswith("BRA"):
        m = re.search(r"0x([0-9a-f]+)", rest)
        if m and int(m.group(1), 16) <= a:
            loops.append((int(m.group(1), 16), a))
fp = [a for a, op, _ in ins if op in ("DADD", "DFMA", "DMUL")]
big = sorted([l for l in loops if sum(1 for x in fp if l[0] <= x <= l[1]) > 50], key=lambda l: l[0] - l[1])
outer = big[0]
inner = [l for l in big[1:] if outer[0] <= l[0] and l[1] <= outer[1]]
inner = inner[0] if inner else None
spins = [l for l in loops if l not in big]
hist = collections.Counter()
for a, op, _ in ins:
    if not (outer[0] <= a <= outer[1]):
        continue
    if any(s[0] <= a <= s[1] for s in spins) and op.startswith("BRA"):
        pass
    w = w_inner if inner and inner[0] <= a <= inner[1] else 1
    hist[op] += w
tot = sum(hist.values())
f64 = hist["DADD"] + hist["DFMA"] + hist["DMUL"]
print(f"outer loop {outer[0]:#x}-{outer[1]:#x}, inner {inner and (hex(inner[0]), hex(inner[1]))}, code bytes {16 * len(ins)}")
print(f"per warp-step: {tot} instructions, {f64} FP64, {tot - f64} other; issue-cycle estimate 2*FP64 + other = {2 * f64 + tot - f64}")
print("  ".join(f"{k}={v}" for k, v in hist.most_common(40)))
