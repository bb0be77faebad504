"""Deterministic synthetic weights and inputs for the temporal head.

No dataset or checkpoint is available offline, so tests and ``bench.py`` use
seeded synthetic tensors of the reference's shapes.  The
reference's own initialisation (AffineDropPath scale 1e-4, Conv2d std 1e-3,
model/blocks.py:289-295, model/OTPose.py:438-439) makes every residual branch
vanish, which would let a parity test pass vacuously; the rules below give
every branch O(1) weight instead.

Values are a pure function of (seed, key name, shape) using numpy's PCG64
stream, so they do not depend on module iteration order or on torch's RNG.
"""
from __future__ import annotations

import math
import zlib

import numpy as np
import torch


def sinusoid_table(n_position: int, d_hid: int) -> torch.Tensor:
    """(1, d_hid, n_position) fp32 sinusoid table, float64 evaluation then cast --
    the ``pos_embd`` buffer contents before the 1/sqrt(C) rescale
    (reference model/blocks.py:114-125, model/ConvVideoTransformer.py:55-58)."""
    pos = np.arange(n_position, dtype=np.float64)[:, None]
    j = np.arange(d_hid)[None, :]
    ang = pos / np.power(10000.0, 2 * (j // 2) / d_hid)
    tab = np.where(j % 2 == 0, np.sin(ang), np.cos(ang))
    return torch.from_numpy(tab.astype(np.float32).T.copy()).unsqueeze(0)


def _rng(seed: int, key: str) -> np.random.Generator:
    return np.random.default_rng([seed, zlib.crc32(key.encode())])


def synth_tensor(key: str, shape, seed: int = 2024, dtype=torch.float32) -> torch.Tensor:
    """Value rule for one state-dict entry, chosen from its (reference) key name."""
    shape = tuple(shape)
    r = _rng(seed, key)
    leaf = key.split(".")[-1]
    parent = key.split(".")[-2] if "." in key else ""

    def normal(std):
        return r.standard_normal(shape).astype(np.float32) * np.float32(std)

    def uniform(lo, hi):
        return r.uniform(lo, hi, shape).astype(np.float32)

    if leaf == "num_batches_tracked":
        return torch.zeros(shape, dtype=torch.int64)
    if leaf == "pos_embd":
        _, c, n = shape
        return (sinusoid_table(n, c) / (c ** 0.5)).to(dtype)
    if leaf == "scale":                                  # AffineDropPath
        v = uniform(0.5, 1.5)
    elif leaf == "running_mean":
        v = normal(0.1)
    elif leaf == "running_var":
        v = uniform(0.5, 1.5)
    elif parent in ("ln1", "ln2", "bn") or parent.endswith("_norm"):
        v = uniform(0.5, 1.5) if leaf == "weight" else normal(0.1)
    elif leaf == "bias":
        v = normal(0.05)
    elif parent.endswith("_conv") and len(shape) == 3 and shape[1] == 1:   # depthwise Conv1d
        b = 1.0 / math.sqrt(shape[2])
        v = uniform(-b, b)
    elif len(shape) == 3:                                # pointwise Conv1d
        b = 1.0 / math.sqrt(shape[1] * shape[2])
        v = uniform(-b, b)
    elif parent == "deform_conv":                        # DCN weight: identity centre tap + noise
        v = normal(0.05)
        co, ci, kh, kw = shape
        for k in range(min(co, ci)):
            v[k, k, kh // 2, kw // 2] += 1.0
    elif key.startswith("offsets_list"):
        v = normal(2.0 / math.sqrt(shape[1] * shape[2] * shape[3]))
    elif len(shape) == 4:                                # Conv2d (RSB, final, mask convs)
        v = normal(1.0 / math.sqrt(shape[1] * shape[2] * shape[3]))
    else:
        v = normal(0.1)
    return torch.from_numpy(np.ascontiguousarray(v)).to(dtype)


def fill_state_dict(shapes: dict, seed: int = 2024) -> dict:
    """``shapes``: key -> shape (e.g. ``{k: v.shape for k, v in module.state_dict().items()}``)."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def synth_rough_heatmaps(batch: int, joints: int, h: int, w: int, frames: int = 5,
                         seed: int = 1234) -> torch.Tensor:
    """(frames*batch, joints, h, w) fp32, frame-major like the backbone output that
    ``OTPose.forward`` splits (reference model/OTPose.py:319-321): one Gaussian
    blob (sigma 3, peak U(0.3,1)) per (clip, frame, joint) with +-2 px jitter
    between frames, plus N(0, 0.01) noise -- the shape of the reference's
    training targets (utils/heatmap.py:48-105, MODEL.SIGMA 3)."""
    r = np.random.default_rng(seed)
    ys = np.arange(h, dtype=np.float32)[:, None]
    xs = np.arange(w, dtype=np.float32)[None, :]
    cx = r.uniform(0, w - 1, (batch, joints))
    cy = r.uniform(0, h - 1, (batch, joints))
    out = np.empty((frames, batch, joints, h, w), dtype=np.float32)
    for f in range(frames):
        jx = cx + r.uniform(-2, 2, (batch, joints))
        jy = cy + r.uniform(-2, 2, (batch, joints))
        peak = r.uniform(0.3, 1.0, (batch, joints)).astype(np.float32)
        g = np.exp(-((xs[None, None] - jx[..., None, None].astype(np.float32)) ** 2 +
                     (ys[None, None] - jy[..., None, None].astype(np.float32)) ** 2) / (2 * 3.0 ** 2))
        out[f] = peak[..., None, None] * g.astype(np.float32)
    out += r.standard_normal(out.shape).astype(np.float32) * np.float32(0.01)
    return torch.from_numpy(out.reshape(frames * batch, joints, h, w))


def synth_margin(batch: int, seed: int = 1236, frames: int = 5) -> torch.Tensor:
    """(batch, frames-1) int64 frame gaps in {0, 1, 2} (reference dataset/PoseTrackDataset.py:263-293;
    4 columns for the reference's 5-frame window)."""
    r = np.random.default_rng(seed)
    return torch.from_numpy(r.integers(0, 3, (batch, frames - 1)).astype(np.int64))


def synth_center_scale(batch: int, seed: int = 1237):
    r = np.random.default_rng(seed)
    center = r.uniform(100, 1000, (batch, 2)).astype(np.float32)
    scale = r.uniform(0.5, 3.0, (batch, 2)).astype(np.float32)
    return center, scale
