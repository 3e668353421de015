#!/usr/bin/env python
"""Per-kernel constants bench.py quotes from an `ncu --set full --import-source on` capture of ITS OWN default
workload: DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum), FP64 pipe utilisation
(sm__pipe_fp64_cycles_active), the FP64 flop actually executed per launch (thread-level, predicated-on
DFMA x 2 + DMUL + DADD from the report's SASS table) and the launch duration under ncu.
usage: python profiles/extract_ncu_constants.py <rep> <workload> <chains> <substeps> <theta kind: posterior|narrow>
merges the entry "<workload>/m<substeps>/<theta kind>" into profiles/ncu_constants.json
The FIRST launch of each kernel role in the report is used (pass-1 ODE, k_mid, pass-2 ODE, k_score, k_accept_propose)."""
import collections
import csv
import io
import json
import re
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = {h: i for i, h in enumerate(hdr)}


def role(name):
    if name.startswith("void k_mid") or "k_mid<" in name:
        return "profiles_threshold"
    if "k_score<" in name:
        return "score"
    if "k_accept_propose" in name:
        return "accept_propose"
    m = re.search(r"k_ode<\d+, (\d), \d, (\d)>", name)
    if m:
        return None if m.group(2) == "1" else ("ode_pass2" if m.group(1) == "1" else "ode_pass1")
    m = re.search(r"k_ode_age[24]<\d+, (\d)", name)
    if m:
        return "ode_pass2" if m.group(1) == "1" else "ode_pass1"
    return None


# thread-level FP64 flop per launch from the source page (sections in launch order, each printed twice)
flop_by_launch = []
sec, cur, h2, nsec = None, None, None, 0
for r in csv.reader(io.StringIO(src)):
    if r and r[0] == "Kernel Name":
        if nsec % 2 == 0:
            cur = collections.Counter()
            flop_by_launch.append((r[1], cur))
        else:
            cur = None
        nsec += 1
        continue
    if r and r[0] == "Address":
        h2 = r
        continue
    if cur is not None and h2 and len(r) > 8:
        op = r[h2.index("Source")].strip().split()
        o = op[1] if op[0].startswith("@") else op[0]
        o = o.split(".")[0]
        cur[o] += int(r[h2.index("Predicated-On Thread Instructions Executed")])

out = {}
for li, r in enumerate(rows[2:]):
    name = r[idx["Kernel Name"]]
    ro = role(name)
    if ro is None or ro in out:
        continue
    ops = flop_by_launch[li][1] if li < len(flop_by_launch) else collections.Counter()
    f = lambda k: float(r[idx[k]]) if k in idx and r[idx[k]] else None
    units = rows[1]
    def bytes_of(k):
        v, u = f(k), units[idx[k]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    dur, du = f("gpu__time_duration.sum"), units[idx["gpu__time_duration.sum"]]
    out[ro] = {
        "kernel": name.split("(")[0][:60],
        "traffic_bytes": bytes_of("dram__bytes_read.sum") + bytes_of("dram__bytes_write.sum"),
        "fp64_pipe_pct": f("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": f("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "executed_fp64_flop": 2 * ops["DFMA"] + ops["DMUL"] + ops["DADD"],
        "shared_wavefronts": f("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
        "registers": f("launch__registers_per_thread"),
        "duration_ms_under_ncu": dur * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[du],
    }
import os
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ncu_constants.json")
allc = json.load(open(path)) if os.path.exists(path) else {}
allc[f"{sys.argv[2]}/m{int(sys.argv[4])}/{sys.argv[5]}"] = {"source": rep.split("/")[-1], "workload": sys.argv[2], "chains": int(sys.argv[3]),
                                                           "substeps": int(sys.argv[4]), "theta": sys.argv[5], "kernels": out}
json.dump(allc, open(path, "w"), indent=1)
print(json.dumps(allc[f"{sys.argv[2]}/m{int(sys.argv[4])}/{sys.argv[5]}"], indent=1))
