"""CPU restatement of the reference's INPUT WINDOW ASSEMBLY -- TEST INFRASTRUCTURE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this module; the product path
(``otpose_b200/dataset/window.py`` -> ``otp_window_assemble``) never does.

What the reference does per person-clip at inference (``dataset/PoseTrackDataset.py:227-451``, eval path: no
augmentation, rotation 0; ``script/Common.py:343-348``):

1. ``_get_spatio_temporal_window`` picks the supplementary frames prev / next / pprev / nnext around the current
   frame and the four ``margin_*`` integers (``PoseTrackDataset.py:227-317``)       -> :func:`frame_window`
2. ``trans = get_affine_transform(center, scale, 0, image_size)`` (``utils/transform.py:76-107``) -> :func:`get_affine_transform`
3. ``cv2.warpAffine(frame, trans, (W, H), flags=cv2.INTER_LINEAR)`` on each of the five uint8 frames
   (``PoseTrackDataset.py:389-399``; optional BGR -> RGB first, ``:330-337``)         -> :func:`warp_affine_u8`
4. ``ToTensor`` + ``Normalize(mean, std)`` (``utils/transform.py:6-17``)              -> :func:`to_tensor_normalize`
5. ``concat_input = cat((x, prev, next, pprev, nnext), 1)``, ``margin = stack([left, right, lleft, rright], 1)``
   (``script/Common.py:343-348``)                                                   -> :func:`assemble_window`

Third-party arithmetic that is NOT under /root/reference (requirements.txt pins ``opencv-python==4.4.0.44``,
``torchvision==0.8.1``): ``cv2.warpAffine`` and ``cv2.getAffineTransform`` (OpenCV ``modules/imgproc/src/imgwarp.cpp``),
``torchvision.transforms.ToTensor / Normalize``.  Their published algorithms are restated below; PINNED by
``tests/golden/window_*.npz``, which ``oracle/make_golden_window.py`` writes by running the reference's own
``PoseTrackDataset._get_spatio_temporal_window`` (unmodified, file reads served from memory) with the cv2 4.13 and
torchvision of the build container -- the fixed-point warp below reproduces it bit for bit.
"""
from __future__ import annotations

import numpy as np

MEAN = (0.485, 0.456, 0.406)   # utils/transform.py:7-8 (RGB)
STD = (0.229, 0.224, 0.225)


def frame_window(current_idx, num_frames, is_posetrack18, distance, exists=lambda idx: True):
    """PoseTrackDataset.py:243-303.  Returns ((prev, next, pprev, nnext) frame numbers, (margin_left, margin_right,
    margin_lleft, margin_rright)).  ``exists(idx)``: whether the frame file is there (``:305-311``: a missing prev /
    next falls back to the current frame with margin 0; pprev / nnext are NOT checked by the reference)."""
    far = distance
    prev_range = list(range(1, min((current_idx + 1) if is_posetrack18 else current_idx, far + 1)))
    next_range = list(range(1, min((num_frames - current_idx) if is_posetrack18 else (num_frames - current_idx + 1),
                                   far + 1)))
    if len(prev_range) == 0:
        prev_delta = margin_left = pprev_delta = margin_lleft = 0
    elif len(prev_range) == 1:
        prev_delta = margin_left = prev_range[0]
        pprev_delta = margin_lleft = 0
    else:
        prev_delta = margin_left = prev_range[0]
        pprev_delta = margin_lleft = prev_range[1]
    if len(next_range) == 0:
        next_delta = margin_right = nnext_delta = margin_rright = 0
    elif len(next_range) == 1:
        next_delta = margin_right = next_range[-1]
        nnext_delta = margin_rright = 0
    else:
        next_delta = margin_right = next_range[0]
        nnext_delta = margin_rright = next_range[0]       # sic (:291): the reference takes [0] again, not [1]
    prev_idx, next_idx = current_idx - prev_delta, current_idx + next_delta
    pprev_idx, nnext_idx = current_idx - pprev_delta, current_idx + nnext_delta
    if not exists(prev_idx):
        prev_idx, margin_left = current_idx, 0
    if not exists(next_idx):
        next_idx, margin_right = current_idx, 0
    return (prev_idx, next_idx, pprev_idx, nnext_idx), (margin_left, margin_right, margin_lleft, margin_rright)


def _get_affine_3pt(src, dst):
    """cv2.getAffineTransform(src, dst): the 2x3 double matrix mapping three float32 points (imgwarp.cpp:
    a 6x6 linear system solved in double)."""
    a = np.zeros((6, 6), np.float64)
    b = np.zeros(6, np.float64)
    for i in range(3):
        a[i, 0:2] = src[i]
        a[i, 2] = 1
        a[i + 3, 3:5] = src[i]
        a[i + 3, 5] = 1
        b[i], b[i + 3] = dst[i, 0], dst[i, 1]
    return np.linalg.solve(a, b).reshape(2, 3)


def get_affine_transform(center, scale, rot, output_size):
    """utils/transform.py:76-107 (shift = 0, inv = 0).  The reference mixes float32 arrays with Python floats, so its
    intermediate precision depends on the NumPy version (1.19, pinned by requirements.txt, promotes to float64; NumPy
    2 keeps float32): the matrix is reproduced to ~1e-7, which is why the warp takes the matrix as an INPUT."""
    scale = np.asarray(scale, np.float64) if isinstance(scale, (list, tuple, np.ndarray)) else np.array([scale, scale])
    scale_tmp = scale * 200.0
    src_w, dst_w, dst_h = scale_tmp[0], output_size[0], output_size[1]
    rot_rad = np.pi * rot / 180
    sn, cs = np.sin(rot_rad), np.cos(rot_rad)
    src_dir = np.array([0 * cs - (src_w * -0.5) * sn, 0 * sn + (src_w * -0.5) * cs])
    dst_dir = np.array([0, dst_w * -0.5], np.float32)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = center
    src[1] = np.asarray(center) + src_dir
    dst[0] = [dst_w * 0.5, dst_h * 0.5]
    dst[1] = np.array([dst_w * 0.5, dst_h * 0.5]) + dst_dir
    third = lambda p, q: q + np.array([-(p - q)[1], (p - q)[0]], np.float32)   # noqa: E731  get_3rd_point
    src[2] = third(src[0], src[1])
    dst[2] = third(dst[0], dst[1])
    return _get_affine_3pt(np.float32(src), np.float32(dst))


def warp_affine_u8(src, trans, dsize):
    """cv2.warpAffine(src (Hs, Ws, C) uint8, trans (2, 3), dsize = (W, H), flags=INTER_LINEAR), BORDER_CONSTANT 0.

    OpenCV imgwarp.cpp: the matrix is inverted in double; per destination pixel the source position is a FIXED-POINT
    number (AB_BITS = 10 scale, rounded to 1/32 pixel with round_delta = 16), the four bilinear weights come from a
    32 x 32 table of 15-bit integers ((32 - fy)(32 - fx) * 32, ...), the result is (sum + 2^14) >> 15."""
    w, h = int(dsize[0]), int(dsize[1])
    m = np.array(trans, np.float64).reshape(6).copy()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    ab_scale = 1 << 10
    rnd = lambda v: np.rint(v).astype(np.int64)   # noqa: E731  cvRound / saturate_cast<int>: half to even
    x = np.arange(w, dtype=np.float64)
    y = np.arange(h, dtype=np.float64)
    adelta, bdelta = rnd(m[0] * x * ab_scale), rnd(m[3] * x * ab_scale)
    x0 = rnd((m[1] * y + m[2]) * ab_scale) + 16
    y0 = rnd((m[4] * y + m[5]) * ab_scale) + 16
    X = (x0[:, None] + adelta[None, :]) >> 5
    Y = (y0[:, None] + bdelta[None, :]) >> 5
    sx, sy = np.clip(X >> 5, -32768, 32767), np.clip(Y >> 5, -32768, 32767)   # saturate_cast<short>
    fx, fy = X & 31, Y & 31
    w00, w01, w10, w11 = (32 - fy) * (32 - fx) * 32, (32 - fy) * fx * 32, fy * (32 - fx) * 32, fy * fx * 32
    sh, sw = src.shape[:2]

    def px(yy, xx):
        ok = (yy >= 0) & (yy < sh) & (xx >= 0) & (xx < sw)
        v = src[np.clip(yy, 0, sh - 1), np.clip(xx, 0, sw - 1)].astype(np.int64)
        return v * ok[..., None]

    acc = px(sy, sx) * w00[..., None] + px(sy, sx + 1) * w01[..., None] + px(sy + 1, sx) * w10[..., None] + \
        px(sy + 1, sx + 1) * w11[..., None]
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def to_tensor_normalize(img_u8):
    """torchvision ToTensor (uint8 HWC -> float32 CHW, / 255) then Normalize: (x - mean) / std, all in fp32
    (utils/transform.py:11-17)."""
    x = img_u8.astype(np.float32).transpose(2, 0, 1) / np.float32(255)
    mean = np.array(MEAN, np.float32)[:, None, None]
    std = np.array(STD, np.float32)[:, None, None]
    return (x - mean) / std


def assemble_window(frames, frame_ids, center, scale, image_size, color_rgb=True, trans=None):
    """One clip: frames = dict / sequence indexable by frame number -> (Hs, Ws, 3) uint8 BGR (cv2.imread order);
    frame_ids = (cur, prev, next, pprev, nnext).  Returns (15, H, W) float32 (PoseTrackDataset.py:330-406 +
    Common.py:347) and the 2x3 transform (``trans``: use this matrix instead of recomputing it)."""
    if trans is None:
        trans = get_affine_transform(center, scale, 0, image_size)
    outs = []
    for fid in frame_ids:
        img = frames[fid]
        if color_rgb:
            img = img[:, :, ::-1]
        outs.append(to_tensor_normalize(warp_affine_u8(img, trans, (int(image_size[0]), int(image_size[1])))))
    return np.concatenate(outs, 0), trans
