#!/usr/bin/env python
"""Generate tests/golden/window_*.npz by RUNNING THE REFERENCE's own input pipeline on CPU.

Executed unmodified: ``dataset/PoseTrackDataset.py`` ``PoseTrackDataset._get_spatio_temporal_window`` (lines 227-451:
frame selection, margins, ``get_affine_transform``, ``cv2.warpAffine``, ``build_transforms``), ``utils/transform.py``,
and the ``torch.cat`` / ``torch.stack`` of ``script/Common.py:343-348``.  Shims (none changes arithmetic on the path):

* absent packages imported at module level (``matplotlib``, ``pycocotools``, ``yacs``, ``motmetrics``, ``shapely``,
  ``tensorboardX``: evaluation / plotting / config only) -> empty stub modules;
* the dataset object is a bare namespace carrying the attributes the method reads (no annotation files);
* ``cv2.imread`` / ``os.path.exists`` inside the dataset module are served from an in-memory frame table (synthetic
  uint8 frames; a missing file is a missing table entry), so no image files are needed.

Runs only in the build container (needs /root/reference + cv2); the tests read the committed vectors.
Usage:  python oracle/make_golden_window.py
"""
from __future__ import annotations

import os
import sys
import types

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("OTPOSE_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
GOLD = os.path.join(REPO, "tests", "golden")


def install_shims():
    for name in ("matplotlib", "matplotlib.pyplot", "pycocotools", "pycocotools.coco", "yacs", "yacs.config", "tensorboardX",
                 "motmetrics", "shapely", "shapely.geometry", "scipy.io"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["yacs.config"].CfgNode = dict


def synth_frames(n, hs, ws, seed):
    """Smooth colour gradients + blobs + noise: every bilinear weight matters, borders are not black."""
    r = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:hs, 0:ws].astype(np.float32)
    out = np.empty((n, hs, ws, 3), np.uint8)
    for f in range(n):
        img = np.stack([40 + 150 * xx / ws, 60 + 120 * yy / hs, 200 - 100 * (xx + yy) / (hs + ws)], -1)
        for _ in range(6):
            cy, cx, s = r.uniform(0, hs), r.uniform(0, ws), r.uniform(4, 20)
            img += r.uniform(-90, 90, 3) * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))[..., None]
        img += r.normal(0, 6, img.shape)
        out[f] = np.clip(img, 0, 255).astype(np.uint8)
    return out


def main():
    install_shims()
    import dataset.PoseTrackDataset as P          # the reference module, unmodified
    from utils.transform import build_transforms

    for name, is18, color_rgb, (hs, ws), image_size, n_frames, seed in (
            ("window_pt17", False, True, (90, 120), (48, 64), 7, 11),
            ("window_pt18", True, False, (72, 100), (36, 48), 6, 12)):
        frames = synth_frames(n_frames, hs, ws, seed)
        zero_fill = 6 if is18 else 8
        first = 0 if is18 else 1                    # PoseTrack18 frame files start at 000000.jpg, PoseTrack17 at 00000001.jpg
        table = {f"/v/{str(first + i).zfill(zero_fill)}.jpg": frames[i] for i in range(n_frames)}
        missing = {f"/v/{str(first + 1).zfill(zero_fill)}.jpg"}      # one supplementary file is "not on disk"
        P.cv2.imread = lambda p: None if p not in table else table[p].copy()
        P.osp.exists = lambda p: p in table and p not in missing
        ds = types.SimpleNamespace(distance=2, color_rgb=color_rgb, train=False, transform=build_transforms(None, "val"),
                                   image_size=np.array(image_size), num_joints=17, sigma=3,
                                   heatmap_size=np.array([image_size[0] // 4, image_size[1] // 4]),
                                   use_different_joints_weight=False, joints_weight=1)
        r = np.random.default_rng(seed + 100)
        clips = []
        for cur in range(n_frames):
            for _ in range(2):
                center = np.array([r.uniform(0.1, 0.9) * ws, r.uniform(0.1, 0.9) * hs], np.float32)
                s = r.uniform(0.15, 0.6)
                scale = np.array([s * image_size[0] / image_size[1], s], np.float32) * 1.25
                item = dict(filename="v", imgnum=cur, image=f"/v/{str(first + cur).zfill(zero_fill)}.jpg",
                            nframes=n_frames, joints_3d=np.zeros((17, 3), np.float32),
                            joints_3d_vis=np.zeros((17, 3), np.float32), center=center.copy(), scale=scale.copy(), score=1)
                x, prev, nxt, pprev, nnext, _, _, meta = P.PoseTrackDataset._get_spatio_temporal_window(ds, item)
                concat = torch.cat((x, prev, nxt, pprev, nnext), 0)                       # Common.py:347 (dim 1 of the batch)
                margin = [meta["margin_left"], meta["margin_right"], meta["margin_lleft"], meta["margin_rright"]]
                clips.append((first + cur, center, scale, concat.numpy(), margin,
                              P.get_affine_transform(center, scale, 0, ds.image_size)))
        np.savez_compressed(
            os.path.join(GOLD, name + ".npz"), frames=frames, first_frame=first, is_posetrack18=int(is18),
            color_rgb=int(color_rgb), distance=2, image_size=np.array(image_size),
            missing=np.array([first + 1]), current=np.array([c[0] for c in clips]),
            center=np.stack([c[1] for c in clips]), scale=np.stack([c[2] for c in clips]),
            concat_input=np.stack([c[3] for c in clips]), margin=np.array([c[4] for c in clips], np.int64),
            trans=np.stack([c[5] for c in clips]))
        print(name, "clips", len(clips), "concat_input", np.stack([c[3] for c in clips]).shape)


if __name__ == "__main__":
    main()
