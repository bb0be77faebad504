"""f4: input window assembly -- frame selection / margins, get_affine_transform, the
cv2.warpAffine + ToTensor + Normalize + concat of the five frames of a clip.

CPU tests: the oracle (oracle/window_oracle.py) and the host logic of the product module against fixtures written by
RUNNING the reference's own ``PoseTrackDataset._get_spatio_temporal_window`` (oracle/make_golden_window.py).
GPU tests: ``otp_window_assemble`` through the C ABI, bit-exact against the fixtures and against the oracle at full
size (384 x 288 crops of 720p frames, crops hanging over the frame border)."""
import os

import numpy as np
import pytest
import torch

from oracle import window_oracle as wo
from otpose_b200.dataset import window as win

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIXTURES = ("window_pt17", "window_pt18")


def fixture(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    frames, first = g["frames"], int(g["first_frame"])
    table = {first + i: frames[i] for i in range(len(frames))}
    missing = set(g["missing"].tolist())
    return g, table, first, (lambda idx: idx in table and idx not in missing)


@pytest.mark.parametrize("name", FIXTURES)
def test_window_oracle_vs_reference_golden(name):
    """The restated warpAffine / ToTensor / Normalize / frame selection reproduce the reference pipeline BIT FOR BIT."""
    g, table, first, exists = fixture(name)
    for i, cur in enumerate(g["current"].tolist()):
        ids, margin = wo.frame_window(cur, len(table), bool(g["is_posetrack18"]), int(g["distance"]), exists)
        out, _ = wo.assemble_window(table, (cur,) + ids, g["center"][i], g["scale"][i], g["image_size"],
                                    bool(g["color_rgb"]), trans=g["trans"][i])
        assert list(margin) == g["margin"][i].tolist()
        assert np.array_equal(out, g["concat_input"][i])
        # the matrix itself: float32 / float64 intermediates of the reference depend on the NumPy version (see oracle)
        assert np.abs(wo.get_affine_transform(g["center"][i], g["scale"][i], 0, g["image_size"]) - g["trans"][i]).max() < 1e-5


@pytest.mark.parametrize("name", FIXTURES)
def test_window_host_logic_vs_reference_golden(name):
    """Product-side host functions (same names as the reference's): frame numbers, margins, the affine matrix."""
    g, table, first, exists = fixture(name)
    for i, cur in enumerate(g["current"].tolist()):
        ids, margin = win.frame_window(cur, len(table), bool(g["is_posetrack18"]), int(g["distance"]), exists)
        assert (ids, margin) == wo.frame_window(cur, len(table), bool(g["is_posetrack18"]), int(g["distance"]), exists)
        assert list(margin) == g["margin"][i].tolist()
        t = win.get_affine_transform(g["center"][i], g["scale"][i], 0, g["image_size"])
        assert np.abs(t - g["trans"][i]).max() < 1e-5
    # every position of short and long videos, both numbering schemes, distances 1..3, with and without missing files
    for is18 in (False, True):
        for n in (1, 2, 3, 4, 7):
            for dist in (1, 2, 3):
                for cur in range(0 if is18 else 1, n if is18 else n + 1):
                    for ex in (lambda i: True, lambda i: i % 3 != 1):
                        assert win.frame_window(cur, n, is18, dist, ex) == wo.frame_window(cur, n, is18, dist, ex)
    # the two ends of a video and a one-frame video
    assert win.frame_window(0, 5, True, 2) == ((0, 1, 0, 1), (0, 1, 0, 1))
    assert win.frame_window(1, 5, False, 2) == ((1, 2, 1, 2), (0, 1, 0, 1))
    assert win.frame_window(5, 5, False, 2) == ((4, 5, 3, 5), (1, 0, 2, 0))
    assert win.frame_window(0, 1, True, 2) == ((0, 0, 0, 0), (0, 0, 0, 0))


def test_window_needs_cuda_tensors():
    with pytest.raises(NotImplementedError):
        win.assemble_windows(torch.zeros((2, 8, 8, 3), dtype=torch.uint8), [[0, 0, 1, 0, 1]], np.eye(2, 3)[None], (8, 8))


def _gpu_clip_inputs(g, table, first, exists):
    fi, tr, mg = [], [], []
    for i, cur in enumerate(g["current"].tolist()):
        ids, margin = win.frame_window(cur, len(table), bool(g["is_posetrack18"]), int(g["distance"]), exists)
        fi.append([f - first for f in (cur,) + ids])
        tr.append(g["trans"][i])
        mg.append(margin)
    return np.array(fi), np.stack(tr), np.array(mg)


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIXTURES)
def test_window_assemble_vs_reference_golden(name):
    """otp_window_assemble == the reference's own pipeline output, bit for bit (fp32 concat_input and margins)."""
    g, table, first, exists = fixture(name)
    fi, tr, mg = _gpu_clip_inputs(g, table, first, exists)
    frames = torch.from_numpy(g["frames"]).cuda()
    out, img16, margin = win.assemble_windows(frames, fi, tr, g["image_size"], bool(g["color_rgb"]), margin=mg,
                                              bf16_nhwc=True)
    assert out.shape == g["concat_input"].shape and out.dtype == torch.float32
    assert np.array_equal(out.cpu().numpy(), g["concat_input"])
    assert margin.dtype == torch.int64 and np.array_equal(margin.cpu().numpy(), g["margin"])
    # the channels-last bf16 batch of OTPose.forward (cat(x.split(3, 1), 0): frame-major) holds the same pixels
    b = out.shape[0]
    ref16 = torch.cat(out.split(3, dim=1), 0).to(torch.bfloat16)
    assert img16.shape == (5 * b, 3) + tuple(out.shape[2:]) and img16.is_contiguous(memory_format=torch.channels_last)
    assert torch.equal(img16, ref16)


@pytest.mark.gpu
@pytest.mark.parametrize("frames_per_clip", [3, 5, 7])
def test_window_assemble_full_size_vs_oracle(frames_per_clip):
    """Bench shape: 384 x 288 crops out of 720p frames, boxes hanging over every border, a degenerate tiny box, several
    persons sharing frames; 3 / 5 / 7 frames per clip (BASELINE config 5)."""
    r = np.random.default_rng(5)
    hs, ws, nfr, b = 720, 1280, 6, 9
    frames = r.integers(0, 256, (nfr, hs, ws, 3), dtype=np.uint8)
    centers = np.array([[640, 360], [5, 5], [1275, 700], [100, 715], [1279, 0], [640, -40], [-30, 360], [900, 200],
                        [300, 500]], np.float32)
    scales = np.array([[1.2, 1.6], [0.6, 0.8], [0.9, 1.2], [2.4, 3.2], [0.3, 0.4], [1.5, 2.0], [1.5, 2.0], [0.02, 0.027],
                       [4.5, 6.0]], np.float32)
    fi = r.integers(0, nfr, (b, frames_per_clip))
    tr = np.stack([win.get_affine_transform(centers[i], scales[i], 0, (288, 384)) for i in range(b)])
    out, _, _ = win.assemble_windows(torch.from_numpy(frames).cuda(), fi, tr, (288, 384), True)
    got = out.cpu().numpy()
    for i in range(b):
        ref = np.concatenate([wo.to_tensor_normalize(wo.warp_affine_u8(frames[f][:, :, ::-1], tr[i], (288, 384)))
                              for f in fi[i]], 0)
        assert np.array_equal(got[i], ref), i
    # empty batch is a no-op
    e, _, _ = win.assemble_windows(torch.from_numpy(frames).cuda(), np.zeros((0, frames_per_clip), np.int64),
                                   np.zeros((0, 2, 3)), (288, 384))
    assert e.shape == (0, 3 * frames_per_clip, 384, 288)
