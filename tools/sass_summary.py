#!/usr/bin/env python
"""profiles/sass_summary.md: per kernel of libwlb200.so, the SASS evidence of the Blackwell paths (B200_PROFILING.md:
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = TMA tensor loads, UBLKCP = cp.async.bulk, HMMA = mma.sync)
and the instruction footprint that decides what a decode-step launch costs.

    python tools/sass_summary.py [whisperlive_b200/libwlb200.so] > profiles/sass_summary.md
"""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "whisperlive_b200/libwlb200.so"
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
name, ops = None, collections.defaultdict(collections.Counter)
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        name = m.group(1)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
    if name and m:
        ops[name][m.group(1)] += 1
dem = subprocess.run(["c++filt"], input="\n".join(ops), capture_output=True, text=True).stdout.splitlines()
rows = []
for mangled, d in zip(ops, dem):
    c = ops[mangled]
    short = re.sub(r"\(.*", "", d).replace("wl::", "").replace("void ", "")
    n = sum(c.values())
    rows.append((short, n * 16 / 1024, sum(v for k, v in c.items() if k.startswith("UTC") and k.endswith("MMA")), c["LDTM"] + c["STTM"],
                 c["UTMALDG"] + c["UTMASTG"], c["UBLKCP"], c["HMMA"], c["SYNCS"]))
print("# SASS summary of libwlb200.so (sm_100a)\n")
print("`cuobjdump -sass` instruction counts per kernel.  `UTC*MMA` = tcgen05.mma, `LDTM` = tcgen05.ld, `UTMALDG` = TMA tensor load, "
      "`UBLKCP` = cp.async.bulk, `HMMA` = mma.sync (warp-level tensor path), `SYNCS` = mbarrier ops.  KB = instructions x 16 B: the "
      "decode-step kernels are launched ~400 times per token and pay for their cold instruction fetch every time.\n")
print("| kernel | SASS KB | UTC*MMA | LDTM/STTM | UTMALDG | UBLKCP | HMMA | SYNCS |\n|---|---:|---:|---:|---:|---:|---:|---:|")
for r in sorted(rows, key=lambda r: r[0]):
    print(f"| `{r[0][:78]}` | {r[1]:.1f} | {r[2]} | {r[3]} | {r[4]} | {r[5]} | {r[6]} | {r[7]} |")
