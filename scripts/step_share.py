"""Kernel share of ONE steady-state head step from an `ncu --metrics gpu__time_duration.sum --csv` launch list
(the step = the launches between two consecutive fusion_sum kernels).

    python scripts/step_share.py gpurun_out/launches.csv > profiles/rNN_step_kernel_share.txt
"""
import collections
import csv
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    for r in csv.reader(l for l in f if l.startswith('"')):
        if r and r[0].isdigit() and "gpu__time_duration" in r[12]:
            rows.append((r[4], float(r[14]) / 1e3))
starts = [i for i, (n, _) in enumerate(rows) if "fusion_sum_kernel" in n]
if len(starts) < 2:
    sys.exit("need at least two fusion_sum launches in the list")
a, b = starts[-2], starts[-1]
step = rows[a:b]
tot = sum(t for _, t in step)
print(f"one steady-state step = {len(step)} launches, sum of kernel durations {tot:.1f} us")
print("(ncu --metrics gpu__time_duration.sum --clock-control none: serialised, cold cache -- shares, not absolutes)")
agg = collections.OrderedDict()
for n, t in step:
    k = re.sub(r"^void ", "", n)
    k = re.sub(r"\(.*$", "", k).replace("otp::<unnamed>::", "").replace("(anonymous namespace)::", "")
    d = agg.setdefault(k, [0.0, 0])
    d[0] += t
    d[1] += 1
for k, (t, c) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f"{t:10.1f} us {c:4d}x {100 * t / tot:6.1f}%  {k[:90]}")
