#!/usr/bin/env python
"""Turn the artefacts of `scripts/gpu_final.sh <tag>` (gpurun_out/) into the tracked evidence under profiles/:

    python scripts/make_profiles.py r02a [--round r02]

  profiles/<round>_bench_1gpu.json, _bench_reference.json     the bench lines as printed on the B200
  profiles/<round>_ncu_launches_fp16_step.csv                 ncu launch list (gpu__time_duration) of the bench command
  profiles/<round>_step_kernel_share.txt                      kernel share of one steady-state step from that list
  profiles/<round>_ncu_summary.md                             `ncu --set full` headline metrics / stall mix per block kernel
  profiles/<round>_traffic.json                               DRAM bytes per launch, keyed by the kernel's source hash
  profiles/<round>_sass_opcodes.txt                           tcgen05 / TMA opcode counts per object file (cuobjdump -sass)
"""
import collections
import csv
import glob
import io
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (kernel_source_sha)

tag = sys.argv[1]
rnd = sys.argv[sys.argv.index("--round") + 1] if "--round" in sys.argv else "r02"
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")


def cp(src, dst):
    if os.path.exists(os.path.join(G, src)):
        shutil.copyfile(os.path.join(G, src), os.path.join(P, dst))
        print("wrote", dst)


cp(f"{tag}_bench_1gpu.json", f"{rnd}_bench_1gpu.json")
cp(f"{tag}_bench_reference.json", f"{rnd}_bench_reference.json")
cp(f"{tag}_launches.csv", f"{rnd}_ncu_launches_fp16_step.csv")
cp(f"{tag}_step_share.txt", f"{rnd}_step_kernel_share.txt")

# ---- ncu --set full summary + traffic record
rep = os.path.join(G, f"{tag}_prof_blocks.ncu-rep")
if os.path.exists(rep):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True,
                         text=True).stdout
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    names = {"tc_back_kernel": "tc_block_back", "tc_front1_kernel": "tc_block_front", "tc_apply_kernel": "tc_block_apply"}
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    traffic = {}
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        for k, name in names.items():
            if k in d.get("Kernel Name", "") and name not in traffic:
                rd = float(d["dram__bytes_read.sum"]) * scale.get(u["dram__bytes_read.sum"], 1)
                wr = float(d["dram__bytes_write.sum"]) * scale.get(u["dram__bytes_write.sum"], 1)
                traffic[name] = {"dram_bytes_per_launch": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                                 "gpu_time_us": float(d["gpu__time_duration.sum"]),
                                 "source_sha": bench.kernel_source_sha(name),
                                 "capture": f"ncu --set full --clock-control none, scripts/prof_block.py 32 6912 1 "
                                            f"(one stem block of the bench shape), gpurun_out/{tag}_prof_blocks.ncu-rep"}
    with open(os.path.join(P, f"{rnd}_traffic.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    print("wrote", f"{rnd}_traffic.json", {k: v["dram_bytes_per_launch"] for k, v in traffic.items()})
    with open(os.path.join(P, f"{rnd}_ncu_summary.md"), "w") as f:
        f.write(f"# {rnd} ncu summaries (B200, `ncu --set full --clock-control none --import-source on`)\n\n"
                "Workload: one stem TransformerBlock of the bench shape (32 clips x 6912 tokens, C = 136, fp16 operands),\n"
                "`python scripts/prof_block.py 32 6912 1 4`, kernels `tc_front1 / gram_project / tc_apply / tc_back`; summary by\n"
                "`scripts/ncu_summary.py` (headline metrics, stall mix, opcode mix, per-barrier-segment SASS breakdown).\n"
                "Times under ncu are cold-cache and serialised: shares, not absolutes -- compare with the bench line.\n\n```\n"
                + out + "```\n")
    print("wrote", f"{rnd}_ncu_summary.md")

# ---- the round-2 kernels outside the C = 136 blocks: flow encoder (one cluster launch), RSB level kernels
for name, what in (("flow", "the C = 17 flow encoder, 32 clips x 6912 tokens, 6 blocks in one cluster launch (`scripts/prof_flow.py`)"),
                   ("rsb", "offset_mask_combine_conv (51 -> 32 -> 32, 18 level launches) at 32 clips x 96 x 72 (`scripts/prof_rsb.py`)")):
    rep2 = os.path.join(G, f"{tag}_prof_{name}.ncu-rep")
    if not os.path.exists(rep2):
        continue
    raw = subprocess.run(["ncu", "-i", rep2, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
            "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    with open(os.path.join(P, f"{rnd}_ncu_{name}.md"), "w") as f:
        f.write(f"# {rnd} ncu `--set full --clock-control none`: {what}\n\n"
                "Times under ncu are cold-cache and serialised; the graph-replay times are in the log below.\n\n```\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(d.get("Kernel Name", "")[:70] + "\n")
            for k in keys:
                if k in d:
                    f.write(f"  {k:86s} {d[k]}\n")
        for lg in (f"{tag}_time_{name}.log",):
            if os.path.exists(os.path.join(G, lg)):
                f.write("\n" + open(os.path.join(G, lg)).read())
        f.write("```\n")
    print("wrote", f"{rnd}_ncu_{name}.md")

# ---- SASS opcode histogram of the built objects (Blackwell-nativeness in-tree)
ops = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "LDGSTS", "SYNCS", "ELECT", "HMMA", "LDSM", "MOVM", "UCGABAR_ARV"]
lines = ["tcgen05 / TMA opcode counts per object (cuobjdump -sass otpose_b200/build/*.o)",
         "UTCHMMA = tcgen05.mma kind::f16, LDTM = tcgen05.ld, UTMALDG / UTMASTG = cp.async.bulk.tensor load / store,",
         "UBLKCP = cp.async.bulk (non-tensor), UTCBAR = tcgen05.commit, LDGSTS = cp.async; HMMA / LDSM / MOVM = mma.sync / ldmatrix /",
         "movmatrix of the narrow-channel kernels (block_flow, rsb_fused), UCGABAR_ARV = cluster barrier arrive", "",
         f"{'object':22s}" + "".join(f"{o:>9s}" for o in ops)]
for obj in sorted(glob.glob(os.path.join(ROOT, "otpose_b200", "build", "*.o"))):
    sass = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cnt = collections.Counter(m.group(1) for m in re.finditer(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", sass, re.M))
    if any(cnt[o] for o in ops):
        lines.append(f"{os.path.basename(obj):22s}" + "".join(f"{cnt[o]:9d}" for o in ops))
with open(os.path.join(P, f"{rnd}_sass_opcodes.txt"), "w") as f:
    f.write("\n".join(lines) + "\n")
print("\n".join(lines))
