#!/usr/bin/env python
"""Per-source-line cost of one kernel: joins the SASS-level table of an .ncu-rep (--import-source on,
--page source --print-source sass) with the line table of the SAME build (nvdisasm -g of the cubin inside
libepimcmc.so; the instruction order is identical).
usage: python profiles/source_hotspots.py <rep> <kernel regex> <mangled-name substring> [top]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kre, mangled = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "epi_mcmc_b200", "csrc", "libepimcmc.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, "epi_kernels.sm_100a.cubin")],
                      capture_output=True, text=True).stdout.splitlines()
# line table of the wanted function: one entry per instruction, in order
lines, cur, inside = [], None, False
for ln in sass:
    if ln.startswith(".text."):
        inside = mangled in ln
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s", ln):
        lines.append(cur)
src = subprocess.run(["ncu", "-i", rep, "--kernel-name", "regex:" + kre, "--page", "source", "--csv",
                      "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, blocks = None, []
for r in rows:
    if r and r[0] == "Kernel Name":
        blocks.append([])
        continue
    if r and r[0] == "Address":
        hdr = r
        continue
    if blocks and hdr and len(r) > 6:
        blocks[-1].append(r)
# several launches may match the regex (template instances share the base name): take the one whose
# instruction count equals the line table's
data = next((b for b in blocks if len(b) == len(lines)), blocks[0])
ix, ism = hdr.index("Instructions Executed"), hdr.index("# Samples")
print(f"# {len(data)} SASS instructions in the report, {len(lines)} in the line table")
n = min(len(data), len(lines))
ins, smp = collections.Counter(), collections.Counter()
for i in range(n):
    ins[lines[i]] += int(data[i][ix])
    smp[lines[i]] += int(data[i][ism])
ti, ts = sum(ins.values()), sum(smp.values())
texts = {}
print(f"# total executed warp instructions {ti}, samples {ts}")
for key, c in ins.most_common(top):
    fn, line = key if key else ("?", 0)
    if fn not in texts:
        pth = os.path.join(ROOT, "epi_mcmc_b200", "csrc", fn)
        texts[fn] = open(pth).read().splitlines() if os.path.exists(pth) else []
    t = texts[fn][line - 1].strip() if 0 < line <= len(texts[fn]) else "?"
    print(f"{fn[:14]:14s}:{line:5d} {100 * c / ti:5.1f}% instr {100 * smp[key] / max(ts, 1):5.1f}% samples | {t[:100]}")
