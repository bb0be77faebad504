"""Input window assembly on the GPU: the per-clip part of the reference's
``PoseTrackDataset._get_spatio_temporal_window`` (``dataset/PoseTrackDataset.py:227-451``, inference path) and the
``torch.cat`` / ``torch.stack`` of ``script/Common.py:343-348``.

* :func:`frame_window`            the supplementary frame numbers and the four ``margin_*`` integers  (``:243-311``)
* :func:`get_affine_transform`    ``utils/transform.py:76-107`` (same name, same arguments)
* :func:`assemble_windows`        ``cv2.warpAffine`` + ``ToTensor`` + ``Normalize`` of the five frames of every clip +
                                  the channel concat, as ONE launch of ``otp_window_assemble`` on uint8 frames that
                                  are already on the device -> ``concat_input (B, 15, H, W)`` fp32 (bit-exact with the
                                  reference pipeline) and / or the bf16 channels-last ``(5B, H, W, 3)`` batch of
                                  ``model/OTPose.py:317``, plus ``margin (B, 4)`` int64.

CUDA tensors only (no CPU fallback, like every other entry point of this package).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from .. import _lib

MEAN = (0.485, 0.456, 0.406)   # utils/transform.py:7-8 (RGB)
STD = (0.229, 0.224, 0.225)


def frame_window(current_idx, num_frames, is_posetrack18, distance, exists=lambda idx: True):
    """``((prev, next, pprev, nnext), (margin_left, margin_right, margin_lleft, margin_rright))`` for the frame numbered
    ``current_idx`` of a video of ``num_frames`` frames (PoseTrack18 numbers frames from 0, PoseTrack17 from 1).
    ``exists(idx)`` = the frame file is there: a missing prev / next falls back to the current frame with margin 0
    (``PoseTrackDataset.py:305-311``).  ``nnext`` repeats ``next`` exactly as the reference does (``:291``)."""
    prev_range = list(range(1, min((current_idx + 1) if is_posetrack18 else current_idx, distance + 1)))
    next_range = list(range(1, min((num_frames - current_idx) if is_posetrack18 else (num_frames - current_idx + 1),
                                   distance + 1)))
    prev_delta = margin_left = prev_range[0] if prev_range else 0
    pprev_delta = margin_lleft = prev_range[1] if len(prev_range) > 1 else 0
    if len(next_range) == 1:
        next_delta = margin_right = next_range[-1]
        nnext_delta = margin_rright = 0
    else:
        next_delta = margin_right = nnext_delta = margin_rright = next_range[0] if next_range else 0
    prev_idx, next_idx = current_idx - prev_delta, current_idx + next_delta
    if not exists(prev_idx):
        prev_idx, margin_left = current_idx, 0
    if not exists(next_idx):
        next_idx, margin_right = current_idx, 0
    return ((prev_idx, next_idx, current_idx - pprev_delta, current_idx + nnext_delta),
            (margin_left, margin_right, margin_lleft, margin_rright))


def get_affine_transform(center, scale, rot, output_size, shift=np.array([0, 0], dtype=np.float32), inv=0):
    """``utils/transform.py:76-107``: the 2x3 float64 matrix that ``cv2.getAffineTransform`` returns for the three
    point pairs (centre, a point half a box width above it rotated by ``rot`` degrees, their right-angle third)."""
    if not isinstance(scale, (np.ndarray, list, tuple)):
        scale = np.array([scale, scale])
    scale_tmp = np.asarray(scale, np.float64) * 200.0
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    p = (0.0, src_w * -0.5)
    src_dir = np.array([p[0] * cs - p[1] * sn, p[0] * sn + p[1] * cs])
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src, dst = np.zeros((3, 2), np.float32), np.zeros((3, 2), np.float32)
    src[0] = np.asarray(center) + scale_tmp * shift
    src[1] = np.asarray(center) + src_dir + scale_tmp * shift
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    for pts in (src, dst):                      # get_3rd_point
        d = pts[0] - pts[1]
        pts[2] = pts[1] + np.array([-d[1], d[0]], np.float32)
    a, b = (dst, src) if inv else (src, dst)
    m = np.zeros((6, 6), np.float64)            # cv2.getAffineTransform: 6 x 6 system in double
    rhs = np.zeros(6, np.float64)
    for i in range(3):
        m[i, 0:2], m[i, 2] = a[i], 1
        m[i + 3, 3:5], m[i + 3, 5] = a[i], 1
        rhs[i], rhs[i + 3] = b[i, 0], b[i, 1]
    return np.linalg.solve(m, rhs).reshape(2, 3)


def assemble_windows(frames, frame_index, trans, image_size, color_rgb=True, margin=None, fp32=True, bf16_nhwc=False,
                     mean=MEAN, std=STD, out=None, out16=None):
    """frames: ``(F, Hs, Ws, 3)`` uint8 CUDA tensor in ``cv2.imread`` (BGR) order; frame_index: ``(B, 5)`` integers --
    rows of ``frames`` for cur, prev, next, pprev, nnext (3 or 7 columns for the config-5 frame windows); trans:
    ``(B, 2, 3)`` float64 (``get_affine_transform(center, scale, 0, image_size)`` per clip); image_size ``(W, H)``.

    ``out`` / ``out16``: preallocated outputs to fill (a CUDA-graph's static input buffer, for instance).

    Returns ``(concat_input (B, 3 * frames, H, W) fp32 | None, images (frames * B, 3, H, W) bf16 channels-last | None,
    margin (B, 4) int64 on the device | None)``."""
    _lib.require_cuda(frames)
    if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3:
        raise ValueError("frames must be a (F, Hs, Ws, 3) uint8 tensor")
    frames = frames.contiguous()
    dev = frames.device
    fi = torch.as_tensor(np.asarray(frame_index), dtype=torch.int32).to(dev).contiguous() \
        if not torch.is_tensor(frame_index) else frame_index.to(device=dev, dtype=torch.int32).contiguous()
    tr = torch.as_tensor(np.asarray(trans, np.float64)).to(dev).contiguous() \
        if not torch.is_tensor(trans) else trans.to(device=dev, dtype=torch.float64).contiguous()
    b, nf = fi.shape
    if tr.shape != (b, 2, 3):
        raise ValueError("trans must be (B, 2, 3)")
    if nf not in (3, 5, 7):
        raise NotImplementedError(f"frame window of {nf} not built (3, 5 or 7)")
    w, h = int(image_size[0]), int(image_size[1])
    if out is not None:
        if out.shape != (b, 3 * nf, h, w) or out.dtype != torch.float32 or not out.is_contiguous() or out.device != dev:
            raise ValueError("out must be a contiguous (B, 3 * frames, H, W) float32 tensor on the frames' device")
    elif fp32:
        out = torch.empty((b, 3 * nf, h, w), dtype=torch.float32, device=dev)
    if out16 is not None:
        if (out16.shape != (nf * b, 3, h, w) or out16.dtype != torch.bfloat16 or out16.device != dev
                or not out16.is_contiguous(memory_format=torch.channels_last)):
            raise ValueError("out16 must be a channels-last (frames * B, 3, H, W) bfloat16 tensor on the frames' device")
    elif bf16_nhwc:
        out16 = torch.empty((nf * b, 3, h, w), dtype=torch.bfloat16, device=dev).contiguous(
            memory_format=torch.channels_last)
    if out is None and out16 is None:
        raise ValueError("nothing to produce")
    lib = _lib.load()
    mean3, std3 = (C.c_float * 3)(*mean), (C.c_float * 3)(*std)
    with torch.cuda.device(dev):
        _lib.check(lib.otp_window_assemble(
            frames.data_ptr(), frames.shape[0], frames.shape[1], frames.shape[2], frames.stride(0), fi.data_ptr(), nf,
            tr.data_ptr(), b, h, w, int(bool(color_rgb)), mean3, std3, out.data_ptr() if out is not None else None,
            out16.data_ptr() if out16 is not None else None, _lib.stream_ptr(dev)), "otp_window_assemble")
    m = None
    if margin is not None:
        m = torch.as_tensor(np.asarray(margin), dtype=torch.int64).to(dev)      # Common.py:346: stack(...).cuda()
    return out, out16, m
