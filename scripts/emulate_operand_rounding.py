"""CPU diagnostic: which GEMM sites of a TransformerBlock carry the 16-bit operand error.

Emulates the kernels' operand rounding (bf16 / fp16 / split-bf16 hi+lo, fp32 accumulation) on the
oracle's arithmetic, site by site, for temporal_encoder1 of the synthetic head at 96x72 and prints the
error of the stem output `s0` and of `final_layer1(stack(e1))` against the un-rounded run.

    python scripts/emulate_operand_rounding.py
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import head_oracle as ho  # noqa: E402
from otpose_b200.utils import synthetic as syn  # noqa: E402

SITES = ("qk", "gram", "v", "proj", "w1", "w2")


def rounder(kind):
    if kind == "bf16":
        return lambda t: t.bfloat16().float()
    if kind == "fp16":
        return lambda t: t.half().float()
    if kind == "split":      # hi + lo bf16 terms: 16 significand bits
        def f(t):
            hi = t.bfloat16().float()
            return hi + (t - hi).bfloat16().float()
        return f
    return lambda t: t


def block(sd, p, x, n_head, stride, rnd):
    """oracle.transformer_block with r[site] applied to both operands of that site's GEMM."""
    B, C, T = x.shape
    hs = C // n_head
    h = ho.layer_norm_ct(x, sd[p + "ln1.weight"], sd[p + "ln1.bias"])
    a = p + "attn."

    def dw_ln(name):
        y = F.conv1d(h, sd[a + f"{name}_conv.weight"], None, stride=stride, padding=1, groups=C)
        return ho.layer_norm_ct(y, sd[a + f"{name}_norm.weight"], sd[a + f"{name}_norm.bias"])

    def pw(y, name, r):
        return F.conv1d(r[0](y), r[1](sd[a + f"{name}.weight"]), sd[a + f"{name}.bias"])

    q = pw(dw_ln("query"), "query", rnd["qk"]).view(B, n_head, hs, -1)
    k = pw(dw_ln("key"), "key", rnd["qk"]).view(B, n_head, hs, -1)
    vn = dw_ln("value")
    att = rnd["gram"][0](q / math.sqrt(hs)) @ rnd["gram"][1](k).transpose(-2, -1)
    att = F.softmax(att, dim=-1)
    # the kernels fold att @ (Wv vn + bv) = (att Wv) vn + att bv and round W_eff = att Wv once
    wv = sd[a + "value.weight"][:, :, 0].view(n_head, hs, C)
    bv = sd[a + "value.bias"].view(n_head, hs)
    weff = att @ wv.unsqueeze(0)                       # (B, nh, hs, C)
    beff = att @ bv.unsqueeze(0).unsqueeze(-1)         # (B, nh, hs, 1)
    out = rnd["v"][1](weff) @ rnd["v"][0](vn).unsqueeze(1) + beff
    out = out.transpose(2, 3).contiguous().view(B, C, -1)
    out = F.conv1d(rnd["proj"][0](out), rnd["proj"][1](sd[a + "proj.weight"]), sd[a + "proj.bias"])
    skip = x if stride == 1 else F.max_pool1d(x, 3, stride=2, padding=1)
    out = skip + sd[p + "drop_path_attn.scale"] * out
    h2 = ho.layer_norm_ct(out, sd[p + "ln2.weight"], sd[p + "ln2.bias"])
    h2 = F.conv1d(rnd["w1"][0](h2), rnd["w1"][1](sd[p + "mlp.0.weight"]), sd[p + "mlp.0.bias"])
    h2 = F.gelu(h2)
    h2 = F.conv1d(rnd["w2"][0](h2), rnd["w2"][1](sd[p + "mlp.3.weight"]), sd[p + "mlp.3.bias"])
    return out + sd[p + "drop_path_mlp.scale"] * h2


def encoder(sd, p, x, rnd):
    outs = []
    for i in range(6):
        x = block(sd, f"{p}stem.{i}.", x, 2, 1, rnd)
    outs.append(x)
    for i in range(2):
        x = block(sd, f"{p}branch.{i}.", x, 2, 2, rnd)
        outs.append(F.interpolate(x, scale_factor=float(2 ** (i + 1)), mode="linear"))
    return outs


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max()), float((a - b).norm() / b.norm())


def main():
    torch.set_num_threads(os.cpu_count())
    from otpose_b200.model import OTPose, default_cfg
    h, w = 96, 72
    shapes = {k: v.shape for k, v in OTPose(default_cfg((h, w))).state_dict().items()}
    sd = syn.fill_state_dict(shapes, seed=2024)
    rough, margin = syn.synth_rough_heatmaps(1, 17, h, w), syn.synth_margin(1)
    _, inter = ho.head_forward(sd, rough, margin, return_intermediates=True)
    p = "temporal_encoder1."
    x = inter["x1"].flatten(2) + sd[p + "pos_embd"][:, :, :h * w]
    N, Bf, H16, SP = rounder("none"), rounder("bf16"), rounder("fp16"), rounder("split")
    ident = {s: (N, N) for s in SITES}

    def run(rnd):
        e = encoder(sd, p, x, rnd)
        y = torch.stack(e, dim=1).view(1, 408, h, w)
        return e[0], F.conv2d(y, sd["final_layer1.weight"], sd["final_layer1.bias"])

    s0, br = run(ident)
    print("check vs oracle s0", rel(s0, inter["e1"][0]))
    allb = {s: (Bf, Bf) for s in SITES}
    # "gram-first" front: Wq / Wk applied in fp32 after the token reduction; only the Gram operands
    # (the normalised depthwise outputs) are rounded -- emulated as rounding q, k themselves
    gf = {**allb, "qk": (N, N)}
    cases = [("all bf16", allb), ("all fp16", {s: (H16, H16) for s in SITES}),
             ("bf16, qk weights exact (act bf16)", {**allb, "qk": (Bf, N)}),
             ("bf16, qk act exact (weights bf16)", {**allb, "qk": (N, Bf)}),
             ("gram-first bf16", gf),
             ("gram-first + W2 weight split", {**gf, "w2": (Bf, SP)}),
             ("gram-first + W2 act split", {**gf, "w2": (SP, Bf)}),
             ("gram-first + W1,W2 weight split", {**gf, "w1": (Bf, SP), "w2": (Bf, SP)}),
             ("gram-first + W1,W2,Wp weight split", {**gf, "w1": (Bf, SP), "w2": (Bf, SP), "proj": (Bf, SP)}),
             ("gram-first + W1,W2,Wp,Weff weight split", {**gf, "w1": (Bf, SP), "w2": (Bf, SP), "proj": (Bf, SP), "v": (Bf, SP)}),
             ("gram-first + all act split (weights bf16)", {**gf, "w1": (SP, Bf), "w2": (SP, Bf), "proj": (SP, Bf), "v": (SP, Bf)}),
             ("gram-first + all both split", {**gf, "w1": (SP, SP), "w2": (SP, SP), "proj": (SP, SP), "v": (SP, SP)}),
             ("gram-first fp16", {**{s: (H16, H16) for s in SITES}, "qk": (N, N)}),
             ]
    for name, rnd in cases:
        a, b = run(rnd)
        print(f"{name:45s} s0 {rel(a, s0)[0]:.2e} / {rel(a, s0)[1]:.2e}   branches {rel(b, br)[0]:.2e} / {rel(b, br)[1]:.2e}",
              flush=True)


if __name__ == "__main__":
    main()
