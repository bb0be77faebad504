"""Summarise an .ncu-rep: headline metrics, stall mix, opcode mix and per-barrier-segment
breakdown of the SASS (used to write profiles/*_ncu_summary.md).

    python scripts/ncu_summary.py report.ncu-rep
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]


def run(*a):
    return subprocess.run(["ncu", "-i", rep, *a], capture_output=True, text=True).stdout


rows = list(csv.reader(io.StringIO(run("--page", "raw", "--csv"))))
hdr = rows[0]
KEYS = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sm__inst_executed_pipe_tc.sum",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_uniform.sum", "l1tex__data_bank_conflicts_pipe_lsu.sum"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(d.get("Kernel Name", "")[:60])
    for k in KEYS:
        if k in d:
            print(f"  {k:70s} {d[k]}")
src = list(csv.reader(io.StringIO(run("--page", "source", "--csv", "--print-source", "sass"))))
h = None
sass = []
for r in src:
    if len(r) > 5 and r[1] == "Source":
        h = r
        continue
    if h and len(r) == len(h):
        sass.append(dict(zip(h, r)))
I = lambda d, k: int(float(d[k] or 0))  # noqa: E731
tot = sum(I(d, "Instructions Executed") for d in sass)
ts = sum(I(d, "# Samples") for d in sass)
print("SASS lines", len(sass), "warp instr", tot, "samples", ts)
stalls = [x for x in h if x.startswith("stall_") and "Not Issued" not in x]
agg = {s: sum(I(d, s) for d in sass) for s in stalls}
sa = sum(agg.values())
print("stalls", {k[6:]: round(v / sa, 3) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]})
op = collections.Counter()
for d in sass:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
    op[m.group(2).split(".")[0] if m else "?"] += I(d, "Instructions Executed")
print("opcodes", [(o, round(c / tot, 3)) for o, c in op.most_common(16)])
seg, cur = [], []
for d in sass:
    cur.append(d)
    if "BAR.SYNC" in d["Source"] or "EXIT" in d["Source"]:
        seg.append(cur)
        cur = []
if cur:
    seg.append(cur)
for i, s in enumerate(seg):
    ie = sum(I(d, "Instructions Executed") for d in s)
    sm = sum(I(d, "# Samples") for d in s)
    if ie / tot < 0.01 and sm / ts < 0.01:
        continue
    o2 = collections.Counter()
    for d in s:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", d["Source"])
        o2[m.group(2).split(".")[0] if m else "?"] += I(d, "Instructions Executed")
    st = collections.Counter()
    for x in stalls:
        st[x[6:]] += sum(I(d, x) for d in s)
    print(f"seg{i:3d} lines={len(s):5d} instr={ie / tot:.3f} samples={sm / ts:.3f}",
          [(o, round(c / ie, 2)) for o, c in o2.most_common(5)],
          [(k, round(v / max(sm, 1), 2)) for k, v in st.most_common(4)])
