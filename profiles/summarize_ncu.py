#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full --import-source on) into the text summary committed here:
per-kernel headline metrics, stall-reason shares and the executed-instruction opcode mix.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/<name>.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
want = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]
print(f"# {rep}\n")
for r in rows[2:]:
    print("==", r[idx["Kernel Name"]][:110], " id", r[idx["ID"]] if "ID" in idx else "")
    for w in want:
        if w in idx:
            print(f"   {w:72s} {r[idx[w]]:>18s} {units[idx[w]]}")
    items, tot = [], 0.0
    for h in hdr:
        if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued"):
            v = float(r[idx[h]] or 0)
            items.append((v, h.replace("smsp__pcsamp_warps_issue_stalled_", "")))
            tot += v
    print("   stall samples: " + ", ".join(f"{n} {100 * v / tot:.1f}%" for v, n in sorted(items, reverse=True)[:8] if tot))
    print()
rows = list(csv.reader(io.StringIO(src)))
kern, data, h2, nsec = None, collections.OrderedDict(), None, 0
for r in rows:
    if r and r[0] == "Kernel Name":
        # the page prints every launch twice (two views of the same table): keep the first of each pair, so the
        # totals equal smsp__inst_executed.sum above
        kern = r[1] if nsec % 2 == 0 else None
        nsec += 1
        if kern:
            data.setdefault(kern, [])
        continue
    if r and r[0] == "Address":
        h2 = r
        continue
    if kern and len(r) > 6:
        data[kern].append(r)
print("# executed warp-instruction mix (all profiled launches of each kernel summed)")
for k, rs in data.items():
    ix, isrc, ismp = h2.index("Instructions Executed"), h2.index("Source"), h2.index("# Samples")
    tot = sum(int(r[ix]) for r in rs)
    stot = sum(int(r[ismp]) for r in rs)
    ops, smp = collections.Counter(), collections.Counter()
    for r in rs:
        op = r[isrc].strip().split()
        o = op[1] if op[0].startswith("@") else op[0]
        o = o.split(".")[0] + (".MOV" if ".MOV" in o else "")
        ops[o] += int(r[ix])
        smp[o] += int(r[ismp])
    print("==", k[:110], " total", tot)
    for o, c in ops.most_common(16):
        print(f"   {o:10s} {c:14d} {100 * c / tot:5.1f}%   stall samples {100 * smp[o] / max(1, stot):5.1f}%")
    print()
