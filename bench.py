#!/usr/bin/env python
"""Benchmark of the OTPose temporal fusion head on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the whole hot path over the step's
synthetic person-clips: rough heat maps (5B,17,96,72) + margin -> fusion
prologue -> 3 ConvTransformer encoders -> pyramid 1x1 convs -> RSB chains ->
5 x (offset conv, mask conv, modulated DCN) -> refined heat maps -> get_final_preds.
Default workload = BASELINE configs[2]: 512 clips per step, split evenly over the ranks
(strong scaling), every rank running its shard as forward calls of 32 clips (the
configs[1] batch); `--total-clips 0 --batch B` = B clips per GPU per step (weak scaling).
One JSON line on stdout (rank 0).  N > 1 is launched by torchrun: one process per
GPU, no data-path collective (NCCL is only the barrier / max-over-ranks plumbing);
`--train` (configs[3]) adds the gradient all-reduce.

`--impl reference` times the reference's own CPU implementation of the path --
the oracle port (reference ConvTransformer/RSB semantics in torch CPU fp32 +
torchvision deform_conv2d + NumPy get_final_preds) -- on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, J = 96, 72, 17          # --heatmap HxW overrides H, W (BASELINE config 5: 128x96)
FRAMES = 5                     # --frames 3|5|7 (config 5 window sweep; the reference has 5 only)
C8 = 8 * J
T = 96 * 72                    # token count the per-clip FLOP constants below are quoted at
METRIC = "temporal_head_person_clips_per_s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32,
                    help="person-clips per forward call (= per GPU per step unless --total-clips is given)")
    ap.add_argument("--total-clips", type=int, default=512,
                    help="BASELINE configs[2] (the default): a FIXED total of clips per step split over the ranks "
                         "(strong scaling: 512 -> 512/256/128/64 per GPU at 1/2/4/8 GPUs), each rank running its shard "
                         "as forward calls of --batch clips (32 = the configs[1] batch).  0 = weak scaling with "
                         "--batch clips per GPU per step")
    ap.add_argument("--train", action="store_true",
                    help="BASELINE configs[3]: training step of the head (fwd + bwd + bucketed NCCL gradient "
                         "all-reduce + clip + AdamW), --batch clips per GPU (default 64 in this mode)")
    ap.add_argument("--train-micro", type=int, default=16, help="clips per micro-batch of a training step")
    ap.add_argument("--no-full-inference", action="store_true",
                    help="skip the BASELINE configs[1] leg (HRNet-W48 backbone + head, N=1 only)")
    ap.add_argument("--roofline-seconds", type=float, default=2.2,
                    help="minimum GPU time of the per-kernel (roofline) pass, so that the sustained peak applies")
    ap.add_argument("--precision", default="auto", choices=["auto", "fp32", "bf16", "fp16"])
    ap.add_argument("--cpu-clips", type=int, default=2, help="clips per CPU-baseline step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cuda-graph", action="store_true",
                    help="launch the ~190 kernels of a step one by one instead of replaying OTPose's captured graph")
    ap.add_argument("--frames", type=int, default=5, choices=[3, 5, 7],
                    help="frame window (config 5 sweep; 5 = the reference's supplement)")
    ap.add_argument("--heatmap", default="96x72", help="heat-map HxW (config 5: 128x96)")
    args = ap.parse_args()
    global H, W, FRAMES
    H, W = (int(v) for v in args.heatmap.lower().split("x"))
    FRAMES = args.frames
    return args


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def shard_seed(base: int, rank: int) -> int:
    """Every rank draws different synthetic clips (independent shards of the batch)."""
    return base + 1000 * rank


def dist_max(value: float, device=None) -> float:
    """Max over ranks of a python float (device tensor for NCCL, CPU for gloo)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ----------------------------------------------------------------------------------------
# algorithmic work per step (formulas stated in DESIGN.md)
# ----------------------------------------------------------------------------------------
def block_tokens(t):
    """(C, T') of every TransformerBlock launch of one head forward."""
    t1, t2 = t // 2, t // 4
    enc = [(C8, t)] * 6 + [(C8, t1), (C8, t2)]
    return enc + enc + [(J, t)] * 6


def algorithmic_work(b, t=T, precision="fp32"):
    """kernel name -> (bound, work per step): FLOPs (2*MAC) for tensor-bound kernels,
    bytes for HBM-bound ones.  In the 16-bit modes the C=136 blocks run the tc_* kernels and
    the block_* (CUDA-core) kernels only see the C=17 flow encoder."""
    blocks = block_tokens(t)
    big = [(c, tt) for c, tt in blocks if c == C8]
    simt = blocks if precision == "fp32" else [(c, tt) for c, tt in blocks if c != C8]
    w = {}
    flops = lambda f, bl: float(sum(f(c, tt) for c, tt in bl)) * b   # noqa: E731
    w["block_front"] = ("tensor", flops(lambda c, tt: (4 * c * c + 2 * c * (c // (2 if c == C8 else 1))) * tt, simt))
    w["block_apply"] = ("tensor", flops(lambda c, tt: 2 * c * c * tt, simt))
    w["block_back"] = ("tensor", flops(lambda c, tt: 18 * c * c * tt, simt))
    w["tc_block_front"] = ("tensor", flops(lambda c, tt: (4 * c * c + 2 * c * (c // 2)) * tt, big))
    w["tc_block_apply"] = ("tensor", flops(lambda c, tt: 2 * c * c * tt, big))
    w["tc_block_back"] = ("tensor", flops(lambda c, tt: 18 * c * c * tt, big))
    rsb = (0.116e9 + 0.868e9) * t / T * b
    offmask = 5 * (1.218e9 + 0.609e9) * t / T * b
    w["conv2d"] = ("tensor", rsb + offmask if precision == "fp32" else rsb)
    # 16-bit modes: the RSB chains run the per-level mma.sync kernels, the C = 17 flow encoder its one cluster launch
    w["rsb_block"] = ("tensor", rsb)
    w["flow_encoder"] = ("tensor", flops(lambda c, tt: 26 * c * c * tt, [(c, tt) for c, tt in blocks if c != C8]))
    w["tc_offset_mask_dcn"] = ("tensor", offmask + 5 * 0.036e9 * t / T * b)
    # HBM-bound: read offsets+masks (459 ch) + x (17) + write/accumulate out (17), fp32, x5 dilations
    w["mdcn_fwd"] = ("hbm", 5.0 * (459 + 17 + 17) * t * 4 * b)
    w["final_preds"] = ("hbm", float(J * t * 4 + J * 28) * b)
    w["fusion_sum"] = ("hbm", float((FRAMES * J + J + 1) * t * 4) * b)
    w["fusion_stack"] = ("hbm", float((FRAMES * J + J + 1 + 2 * C8 + 2 * J) * t * 4) * b)
    w["pyramid_conv1x1"] = ("hbm", 2.0 * (C8 * (t + t // 2 + t // 4) + J * t) * 4 * b)
    return w


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": p["hbm_gbs"], "tensor": p["bf16_tflops_sustained"], "source": "measured"}
    except Exception:
        return {"hbm": 6650.0, "tensor": 1400.0, "source": "fallback"}


# ----------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md recipe)
# ----------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 <= ts <= t1 and len(r) >= 7] or \
               [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]

        def num(x):
            try:
                return float(x)
            except ValueError:
                return float("nan")
        return {"sm_mhz": statistics.median(num(r[0]) for r in rows), "sm_max_mhz": num(rows[0][1]),
                "power_w_max": max(num(r[2]) for r in rows), "samples": len(rows), "reasons": reasons}


# ----------------------------------------------------------------------------------------
# CPU reference arm / cpu_baseline (the ONLY place bench.py touches oracle/)
# ----------------------------------------------------------------------------------------
def cpu_reference(clips, steps, warmup):
    import torch
    from oracle import head_oracle as ho
    from otpose_b200.model import OTPose, default_cfg
    from otpose_b200.utils import synthetic as syn
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    shapes = {k: v.shape for k, v in OTPose(default_cfg((H, W))).state_dict().items()}
    sd = syn.fill_state_dict(shapes, seed=2024)
    rough = syn.synth_rough_heatmaps(clips, J, H, W, frames=FRAMES)
    margin = syn.synth_margin(clips, frames=FRAMES)
    center, scale = syn.synth_center_scale(clips)

    def step():
        with torch.no_grad():
            out = ho.head_forward(sd, rough, margin)[0]
        ho.get_final_preds(out.numpy(), center, scale)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": clips * steps / dt, "unit": "clips/s", "cores": threads, "kind": "port",
            "sample": f"{steps} steps x {clips} clips of the bench workload ({FRAMES} frames, {H}x{W}, 17 joints), torch CPU fp32 "
                      f"oracle port, {threads} threads", "ms_per_step": 1e3 * dt / steps}


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    cb = cpu_reference(args.cpu_clips, max(1, min(args.steps, 5)), min(args.warmup, 1))
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "clips/s", "n_gpus": args.gpus,
            "steps": max(1, min(args.steps, 5)), "warmup": min(args.warmup, 1), "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong" if args.total_clips else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": dict(workload_config(args.total_clips // max(args.gpus, 1) if args.total_clips else args.batch,
                                           "fp32", args.total_clips, args.gpus, min(args.batch, 32)),
                           sample=f"each timed step is a bounded sample of that workload: {args.cpu_clips} clips on "
                                  f"the host cores (per-clip work identical)"),
            "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": cb["value"], "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def workload_config(batch, precision, total=0, world=1, call=None):
    shard = f"{total} person-clips per step split over {world} GPU(s) = {batch} per GPU, forward calls of {call} clips" \
        if total else f"{batch} person-clips/GPU"
    return {"workload": f"OTPose temporal head fwd (fusion prologue + flow/temporal ConvTransformer encoders + "
                        f"RSB + 5x offset/mask conv + modulated DCN + get_final_preds), {shard}, "
                        f"{FRAMES} frames, {H}x{W} heat maps, {J} joints "
                        f"({'BASELINE configs[2]' if total else 'BASELINE configs[1] batch'}; the HRNet backbone is "
                        f"timed separately in `full_inference`)",
            "clips_per_gpu": batch, "heatmap": [H, W], "joints": J, "precision": precision,
            "l2": "per-step working set (activations > 1 GB at 32 clips) far exceeds the 126 MB L2; no explicit flush"}


# ----------------------------------------------------------------------------------------
def kernel_source_sha(kernel):
    """sha256 (12 hex digits) of the CUDA sources a kernel of the per-kernel table is built from: an ncu
    traffic record (profiles/r02_traffic.json) is only quoted while it matches the code that is running."""
    import hashlib
    files = {"tc_block_back": ["block_tc_back.cuh", "block_tc.cu", "tc_common.cuh"],
             "tc_block_front": ["block_tc_front.cuh", "block_tc.cu", "tc_common.cuh"],
             "tc_block_apply": ["block_tc.cu", "tc_common.cuh"],
             "tc_offset_mask_dcn": ["dcn_tc.cu", "tc_common.cuh"], "conv2d": ["conv_tc.cu", "conv2d.cu", "tc_common.cuh"]}
    h = hashlib.sha256()
    for f in files.get(kernel, []):
        try:
            with open(os.path.join(ROOT, "otpose_b200", "csrc", f), "rb") as fh:
                h.update(fh.read())
        except OSError:
            return None
    return h.hexdigest()[:12] if kernel in files else None


def measured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the round's committed `ncu --set full` capture, or None when
    there is no record for the code as it is now (the record carries the source hash it was taken on)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            rec = json.load(f).get(kernel)
    except Exception:
        return None
    if not rec or rec.get("source_sha") != kernel_source_sha(kernel):
        return None
    return rec.get("dram_bytes_per_launch")


def full_inference(args, model, dev, precision):
    """BASELINE configs[1]: HRNet-W48 backbone (torch / cuDNN, Conv+BN folded, channels-last bf16, CUDA graph)
    -> OTPose.forward_from_features (HRNet.final_layer fused into the head's first kernel) -> final_preds_cuda,
    32 clips x 5 frames of 384x288 synthetic images, random-init weights.  The backbone is library code (SURVEY
    8f rank 3); the leg exists to state the head's share of the full step."""
    import torch
    from otpose_b200.model.HRNet import HRNet, hrnet_w48_cfg
    from otpose_b200.utils import heatmap, synthetic as syn
    b = min(args.batch, 32)
    torch.backends.cudnn.benchmark = True
    net = HRNet(hrnet_w48_cfg(J))
    net.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in net.state_dict().items()}, seed=2025))
    net = net.eval().to(dev)
    fl_w, fl_b = net.final_layer.weight.detach().float(), net.final_layer.bias.detach().float()
    folded = net.fold(dtype=torch.bfloat16, memory_format=torch.channels_last)
    # the step starts from uint8 video frames on the device: the window assembly kernel writes the
    # backbone graph's static bf16 channels-last input
    import numpy as np
    from otpose_b200.dataset import window as win
    rr = np.random.default_rng(1235)
    nfr, hs, ws = 8, 720, 1280
    frames = torch.from_numpy(rr.integers(0, 256, (nfr, hs, ws, 3), dtype=np.uint8)).to(dev)
    sc = rr.uniform(1.0, 2.6, b).astype(np.float32)
    tr = torch.from_numpy(np.stack([win.get_affine_transform(
        np.array([rr.uniform(200, ws - 200), rr.uniform(150, hs - 150)], np.float32), np.array([s * 0.75, s], np.float32), 0,
        (4 * W, 4 * H)) for s in sc])).to(dev)
    fidx = torch.from_numpy(rr.integers(0, nfr, (b, FRAMES))).to(dev)
    images = torch.empty((FRAMES * b, 3, 4 * H, 4 * W), dtype=torch.bfloat16, device=dev) \
        .contiguous(memory_format=torch.channels_last)
    window = lambda: win.assemble_windows(frames, fidx, tr, (4 * W, 4 * H), True, fp32=False, out16=images)   # noqa: E731
    window()
    margin = syn.synth_margin(b, frames=FRAMES).to(dev)
    center, scale = (torch.from_numpy(a).to(dev) for a in syn.synth_center_scale(b))
    with torch.no_grad():
        for _ in range(3):
            feats = folded.features(images)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            feats = folded.features(images)

    def head(f):
        out = model.forward_from_features(f, margin, weight=fl_w, bias=fl_b)[0]
        return heatmap.final_preds_cuda(out, center, scale)

    def step():
        window()
        graph.replay()
        return head(feats)

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    for _ in range(3):
        r = step()
    n = max(3, min(args.steps, 10))
    ms_full = timed(step, n)
    ms_backbone = timed(graph.replay, n)
    ms_head = timed(lambda: head(feats), n)
    ms_window = timed(window, n)
    assert bool(torch.isfinite(r["preds"]).all())
    return {"config": f"full OTPose inference: HRNet-W48 backbone (cuDNN, Conv+BN folded, channels-last bf16, CUDA "
                      f"graph; random init) + temporal head ({precision} operands, HRNet.final_layer fused into the "
                      f"head) + get_final_preds, {b} clips x {FRAMES} frames cropped to {4 * H}x{4 * W} out of {hs}x{ws} uint8 "
                      f"synthetic video frames on the device by the window-assembly kernel, 1 GPU",
            "clips_per_s": round(b / (ms_full * 1e-3), 1), "ms_per_step": round(ms_full, 3),
            "window_ms": round(ms_window, 3), "backbone_ms": round(ms_backbone, 3), "head_ms": round(ms_head, 3),
            "head_share_of_step": round(ms_head / (ms_backbone + ms_head), 4), "steps": n,
            "backbone_tflops": round(353.07e9 * b * (H * W) / T / (ms_backbone * 1e-3) / 1e12, 1),
            "note": "backbone = library (cuDNN) code, BASELINE.md section 3: 353 GFLOP/clip; head = this library"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from otpose_b200 import _lib
    from otpose_b200.model import OTPose, default_cfg
    from otpose_b200.utils import heatmap, synthetic as syn

    rank, local_rank, world = dist_env()
    lib = _lib.load()           # raises when the CUDA library is missing -- no fallback path exists
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    precision = args.precision
    if precision == "auto":
        precision = "fp16" if lib.otp_has_tensor_core_path() else "fp32"
    # clips of this rank per step, and how they are cut into forward calls
    if args.total_clips:
        if args.total_clips % world:
            raise SystemExit(f"--total-clips {args.total_clips} is not divisible by {world} ranks")
        b = args.total_clips // world
    else:
        b = args.batch
    call = min(args.batch, b)
    sizes = [call] * (b // call) + ([b % call] if b % call else [])

    model = OTPose(default_cfg((H, W)), precision=precision, cuda_graph=not args.no_cuda_graph)
    model.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=2024))
    model = model.to(dev).eval()
    model.graph_clone_outputs = False   # the step consumes the outputs at once (final_preds): no defensive copies
    # every forward call of a step has its own synthetic clips (seeded per rank and call), resident on the
    # device for `value`, in pinned host memory for `e2e`
    hosts, resident, cs = [], [], []
    for ci, n in enumerate(sizes):
        sd = 37 * ci
        hosts.append((syn.synth_rough_heatmaps(n, J, H, W, frames=FRAMES, seed=shard_seed(1234 + sd, rank)).pin_memory(),
                      syn.synth_margin(n, seed=shard_seed(1236 + sd, rank), frames=FRAMES).pin_memory()))
        resident.append(tuple(t.to(dev) for t in hosts[-1]))
        cs.append(tuple(torch.from_numpy(a).to(dev) for a in syn.synth_center_scale(n, seed=shard_seed(1237 + sd, rank))))
    preds_host = [torch.empty((n, J, 2), dtype=torch.float32).pin_memory() for n in sizes]
    vals_host = [torch.empty((n, J, 1), dtype=torch.float32).pin_memory() for n in sizes]

    def step_resident():
        for (rough, margin), (center, scale) in zip(resident, cs):
            out = model.forward_head(rough, margin)[0]
            r = heatmap.final_preds_cuda(out, center, scale)
        return r

    # end to end through the public API: every forward call copies ITS inputs from pinned host memory and
    # reads its result back.  Double-buffered: the copy of call i+1 runs on a copy stream while call i
    # computes (forward_head is asynchronous on the caller's stream), as a serving loop would do.
    copy_stream = torch.cuda.Stream(device=dev)
    big = max(range(len(sizes)), key=lambda k: sizes[k])
    stages = [tuple(torch.empty_like(t) for t in resident[big]) for _ in range(2)]
    ev_ready = [torch.cuda.Event(), torch.cuda.Event()]
    ev_free = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0}
    ncall = len(sizes)

    def issue_copy(slot, ci):
        n = sizes[ci]
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[slot])      # the call that last used this slot has consumed it
            stages[slot][0][:FRAMES * n].copy_(hosts[ci][0], non_blocking=True)
            stages[slot][1][:n].copy_(hosts[ci][1], non_blocking=True)
            ev_ready[slot].record(copy_stream)

    def step_e2e():
        main = torch.cuda.current_stream(dev)
        for ci in range(ncall):
            i = e2e_state["i"]
            slot = i & 1
            if i == 0:
                ev_free[0].record(main)
                ev_free[1].record(main)
                issue_copy(0, ci)
            issue_copy(slot ^ 1, (ci + 1) % ncall)     # next call's inputs, overlapped with this call
            main.wait_event(ev_ready[slot])
            n = sizes[ci]
            out = model.forward_head(stages[slot][0][:FRAMES * n], stages[slot][1][:n])[0]
            ev_free[slot].record(main)
            r = heatmap.final_preds_cuda(out, cs[ci][0], cs[ci][1])
            preds_host[ci].copy_(r["preds"], non_blocking=True)
            vals_host[ci].copy_(r["maxvals"], non_blocking=True)
            e2e_state["i"] = i + 1

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        barrier()
        return dist_max(e0.elapsed_time(e1), dev)

    for _ in range(max(args.warmup, 3)):
        step_resident()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    launches0 = lib.otp_launch_count()
    ms = timed(step_resident, args.steps)          # the headline: no per-kernel events in the timed region
    launches = lib.otp_launch_count() - launches0
    # per-kernel table / roofline: a second pass with the library's CUDA events around every launch (2 events
    # x ~190 launches per step cost ~7 % of the step, hence separate), run for at least --roofline-seconds
    # of GPU time so that the SUSTAINED tensor peak is the right denominator
    graph_mode = model.cuda_graph
    model.cuda_graph = False                       # per-kernel events need the individual launches,
    model.overlap_branches = False                 # one after the other (no side-stream overlap)
    lib.otp_profile_enable(1)
    launches1 = lib.otp_launch_count()
    psteps = max(args.steps, min(2000, int(args.roofline_seconds * 1e3 / max(ms / args.steps, 1e-3)) + 1))
    ms_profiled = timed(step_resident, psteps)
    if graph_mode:   # the timed region replayed these same kernels as graph nodes: count them here
        launches = (lib.otp_launch_count() - launches1) * args.steps // psteps
    model.cuda_graph = graph_mode
    model.overlap_branches = True
    prof = _lib.profile_read()
    lib.otp_profile_enable(0)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None

    if rank == 0:
        peaks = load_peaks()
        work = algorithmic_work(b, t=H * W, precision=precision)
        kernels = {}
        for name, (tot_ms, cnt) in prof.items():
            per_step = tot_ms / psteps
            k = {"ms_per_step": round(per_step, 4), "launches_per_step": cnt / psteps}
            if name in work and per_step > 0:
                bound, amount = work[name]
                ach = amount / (per_step * 1e-3) / (1e12 if bound == "tensor" else 1e9)
                k.update(bound=bound, achieved=round(ach, 2), unit="TFLOP/s" if bound == "tensor" else "GB/s",
                         frac=round(ach / peaks[bound], 4))
            kernels[name] = k
        top = max((n for n in kernels if "bound" in kernels[n]), key=lambda n: kernels[n]["ms_per_step"])
        tk = kernels[top]
        roofline = {"kernel": top, "bound": tk["bound"], "achieved": tk["achieved"], "peak": peaks[tk["bound"]],
                    "unit": tk["unit"], "frac": tk["frac"], "traffic": measured_traffic(top),
                    "traffic_source": "profiles/r02_traffic.json (ncu --set full, keyed by the kernel's source hash "
                                      f"{kernel_source_sha(top)}; null = no capture of the code as built)",
                    "peak_source": f"{peaks['source']} ({'sustained bf16' if tk['bound'] == 'tensor' else 'copy'})",
                    "share_of_step": round(tk["ms_per_step"] / sum(k["ms_per_step"] for k in kernels.values()), 4),
                    "timing": "CUDA events around every launch of this kernel in a separate pass of "
                              f"{psteps} steps = {ms_profiled * 1e-3:.2f} s ({ms_profiled / psteps:.3f} ms/step with events)"}
        total = b * world
        line = {"metric": METRIC, "value": total * args.steps / (ms * 1e-3), "unit": "clips/s",
                "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong" if args.total_clips else "weak", "vs_baseline": None,
                "dtype": {"bf16": "bf16", "fp16": "f16", "fp32": "f32"}[precision], "data": "synthetic",
                "config": dict(workload_config(b, precision, args.total_clips, world, call),
                               launch="OTPose(cuda_graph=True): one captured CUDA graph replay per forward call"
                               if model.cuda_graph else "eager: one launch per kernel"), "clocks": clocks,
                "e2e": {"value": total * args.steps / (ms_e2e * 1e-3), "unit": "clips/s",
                        "h2d_bytes_per_step": sum(h[0].numel() * 4 + h[1].numel() * 8 for h in hosts),
                        "d2h_bytes_per_step": sum(p.numel() * 4 for p in preds_host) + sum(v.numel() * 4 for v in vals_host),
                        "ms_per_step": ms_e2e / args.steps,
                        "pipeline": "double-buffered: call i+1's pinned-host -> device copy runs on a copy stream "
                                    "while call i computes; every call's copy and result read-back are inside "
                                    "the timed region"},
                "gpu_launches": int(launches), "roofline": roofline, "kernels": kernels}
        if world == 1 and not args.no_full_inference and (H, W) == (96, 72) and FRAMES == 5:
            try:
                line["full_inference"] = full_inference(args, model, dev, precision)
            except Exception as e:      # the secondary leg must never cost the headline line
                line["full_inference"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_full_inference and (H, W) == (96, 72) and FRAMES == 5:
            try:
                line["window_assembly"] = window_assembly(args, dev)
            except Exception as e:
                line["window_assembly"] = {"unavailable": f"{type(e).__name__}: {e}"[:300]}
        if world == 1 and not args.no_cpu_baseline:
            cb = cpu_reference(args.cpu_clips, 3, 1)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def window_assembly(args, dev):
    """the input window assembly (cv2.warpAffine + ToTensor + Normalize + concat of the five frames
    of every clip) as one launch on uint8 frames resident on the device: 32 clips x 5 frames, 384x288 crops out of
    720p frames -> concat_input (32, 15, 384, 288) fp32.  HBM-bound; algorithmic bytes = the fp32 output + the uint8
    source boxes.  CPU line beside it: the reference's own calls (cv2.warpAffine + torchvision transforms) when cv2
    is importable on the box, else the NumPy oracle port."""
    import time
    import numpy as np
    import torch
    from otpose_b200.dataset import window as win
    b, nfr, hs, ws = min(args.batch, 32), 8, 720, 1280
    r = np.random.default_rng(4321)
    frames = torch.from_numpy(r.integers(0, 256, (nfr, hs, ws, 3), dtype=np.uint8)).to(dev)
    centers = np.stack([r.uniform(200, ws - 200, b), r.uniform(150, hs - 150, b)], 1).astype(np.float32)
    sc = r.uniform(1.0, 2.6, b).astype(np.float32)                       # box height 200 .. 520 px
    scales = np.stack([sc * 0.75, sc], 1)
    fi = r.integers(0, nfr, (b, 5))
    tr = np.stack([win.get_affine_transform(centers[i], scales[i], 0, (4 * W, 4 * H)) for i in range(b)])
    fi_d, tr_d = torch.from_numpy(fi).to(dev), torch.from_numpy(tr).to(dev)
    for _ in range(3):
        out, _, _ = win.assemble_windows(frames, fi_d, tr_d, (4 * W, 4 * H), True)
    # timed through the C ABI on preallocated buffers (the Python wrapper allocates the 212 MB output per call)
    import ctypes as C
    from otpose_b200 import _lib
    lib, n = _lib.load(), 50
    fi32 = fi_d.to(torch.int32).contiguous()
    mean3, std3 = (C.c_float * 3)(*win.MEAN), (C.c_float * 3)(*win.STD)
    call = lambda: _lib.check(lib.otp_window_assemble(   # noqa: E731
        frames.data_ptr(), nfr, hs, ws, frames.stride(0), fi32.data_ptr(), 5, tr_d.data_ptr(), b, 4 * H, 4 * W, 1, mean3,
        std3, out.data_ptr(), None, _lib.stream_ptr(dev)), "otp_window_assemble")
    call()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        call()
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / n
    out_bytes = out.numel() * 4
    src_bytes = float(sum(5 * 3 * (200 * s[0]) * (200 * s[1]) for s in scales))
    peaks = load_peaks()
    res = {"config": f"{b} clips x 5 frames, {4 * H}x{4 * W} crops of {hs}x{ws} uint8 frames on the device -> "
                     f"concat_input ({b},15,{4 * H},{4 * W}) fp32 (bit-exact with cv2.warpAffine + torchvision)",
           "clips_per_s": b / (ms / 1e3), "ms_per_step": ms, "bound": "hbm",
           "achieved": (out_bytes + src_bytes) / (ms / 1e3) / 1e9, "peak": peaks["hbm"], "unit": "GB/s",
           "wire_bytes_per_clip": {"this (48 B transform + 20 B frame ids; frames shared)": 68, "reference (fp32 crops)": out_bytes // b}}
    res["frac"] = res["achieved"] / res["peak"]
    if not args.no_cpu_baseline:
        nb = 4
        try:
            import cv2
            import torchvision.transforms as T
            tf = T.Compose([T.ToTensor(), T.Normalize(mean=win.MEAN, std=win.STD)])
            fr = frames.cpu().numpy()
            t0 = time.time()
            for i in range(nb):
                torch.cat([tf(cv2.warpAffine(cv2.cvtColor(fr[f], cv2.COLOR_BGR2RGB), tr[i], (4 * W, 4 * H), flags=cv2.INTER_LINEAR))
                           for f in fi[i]], 0)
            kind = "reference"
        except ImportError:
            from oracle import window_oracle as wo
            fr = frames.cpu().numpy()
            t0 = time.time()
            for i in range(nb):
                np.concatenate([wo.to_tensor_normalize(wo.warp_affine_u8(fr[f][:, :, ::-1], tr[i], (4 * W, 4 * H))) for f in fi[i]], 0)
            kind = "port"
        res["cpu_baseline"] = {"value": nb / (time.time() - t0), "unit": "clips/s", "cores": 1, "kind": kind,
                               "sample": f"{nb} clips of the same workload, one host thread (a DataLoader worker)"}
    return res


def run_train(args):
    """configs[3]: one training step per rank on its own clips, gradients all-reduced over NCCL.  The JSON
    line states which parts of the step run on this library's kernels and which on library (ATen) ops --
    see otpose_b200/model/train_ops.py."""
    import torch
    import torch.distributed as dist
    from otpose_b200 import _lib
    from otpose_b200.model import OTPose, default_cfg
    from otpose_b200.model.loss import ST_OHKW_MSELoss
    from otpose_b200.train import BucketedGradReducer, train_step
    from otpose_b200.utils import synthetic as syn

    rank, local_rank, world = dist_env()
    lib = _lib.load()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    b = args.batch if args.batch != 32 else 64
    model = OTPose(default_cfg((H, W)))
    model.load_state_dict(syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=2024))
    model = model.to(dev).train()
    params = [p for p in model.parameters() if p.requires_grad]
    red = BucketedGradReducer(params, bucket_bytes=4 << 20)
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True)       # configs/Base_PoseTrack17.yaml:108-109
    crit = ST_OHKW_MSELoss(use_target_weight=True)
    rough = syn.synth_rough_heatmaps(b, J, H, W, seed=shard_seed(1234, rank)).to(dev)
    margin = syn.synth_margin(b, seed=shard_seed(1236, rank)).to(dev)
    target = syn.synth_rough_heatmaps(b, J, H, W, frames=1, seed=shard_seed(1238, rank)).to(dev)
    tw = torch.ones(b, J, 1, device=dev)
    comm_ms = []

    def step():
        return train_step(model, crit, opt, red, rough, margin, target, tw, clip_grad_l2norm=1.0,
                          micro_batch=args.train_micro)

    def barrier():
        if world > 1:
            dist.barrier()

    for _ in range(max(args.warmup, 1)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    wall0 = time.time()
    launches0 = lib.otp_launch_count()
    barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss, norm = step()
    e1.record()
    torch.cuda.synchronize()
    barrier()
    ms = dist_max(e0.elapsed_time(e1), dev)
    launches = lib.otp_launch_count() - launches0
    # the all-reduce alone (same buckets, nothing to overlap with): its share of the step if it were exposed
    if world > 1:
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(10):
            red.zero_grad()
            red.finish()
        c1.record()
        torch.cuda.synchronize()
        comm_ms.append(c0.elapsed_time(c1) / 10)
    wall1 = time.time()
    clocks = sampler.stop(wall0, wall1) if sampler else None
    if rank == 0:
        line = {"metric": "temporal_head_train_person_clips_per_s", "value": b * world * args.steps / (ms * 1e-3),
                "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 1),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": f"OTPose temporal-head training step (BASELINE configs[3]): fwd + bwd + bucketed "
                                       f"NCCL gradient all-reduce + global-norm clip + AdamW, {b} clips/GPU in "
                                       f"micro-batches of {args.train_micro}, {FRAMES} frames, {H}x{W}, {J} joints, "
                                       f"random-init weights, backbone frozen (FREEZE_HRNET_WEIGHTS)",
                           "native": "modulated DCN, offset / mask convs, every RSB conv and the pyramid 1x1 convs: forward and backward on C ABI kernels",
                           "library": "TransformerBlocks, BatchNorm + ReLU of the RSB chains, prologue, loss: ATen ops "
                                      "under autograd (otpose_b200/model/train_ops.py)",
                           "clips_per_gpu": b, "heatmap": [H, W], "joints": J},
                "clocks": clocks, "gpu_launches": int(launches),
                "loss": float(loss), "grad_norm": float(norm),
                "allreduce": {"buckets": red.num_buckets, "bytes": int(red.flat.numel() * 4), "ranks": world,
                              "exposed_ms": comm_ms[0] if comm_ms else 0.0,
                              "share_of_step_if_exposed": (comm_ms[0] / (ms / args.steps)) if comm_ms else 0.0,
                              "overlap": "buckets launched from autograd hooks under the last micro-batch's backward"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.train:
        run_train(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
