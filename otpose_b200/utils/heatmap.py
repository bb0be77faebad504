"""Drop-in ``get_max_preds`` / ``get_final_preds`` (reference utils/heatmap.py:108-171).

The NumPy-signature functions keep the reference contract (host arrays in, host
arrays out) but run the argmax / quarter-pixel / back-projection on the GPU via
``otp_final_preds``; ``final_preds_cuda`` is the device-resident variant that
avoids the full-heatmap D2H copy of script/Common.py:424 (only (N, J, 3) floats
leave the device).
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib


def final_preds_cuda(batch_heatmaps: torch.Tensor, center=None, scale=None):
    """batch_heatmaps (N, J, H, W) fp32 CUDA; center / scale (N, 2) CUDA fp32 or None.

    Returns dict(idx (N,J) int32, coords (N,J,2), preds (N,J,2) or None, maxvals (N,J,1)).
    """
    hm = batch_heatmaps
    assert hm.dim() == 4, 'batch_images should be 4-ndim'
    n, j, h, w = hm.shape
    dev = hm.device
    idx = torch.empty((n, j), dtype=torch.int32, device=dev)
    coords = torch.empty((n, j, 2), dtype=torch.float32, device=dev)
    maxvals = torch.empty((n, j, 1), dtype=torch.float32, device=dev)
    preds = torch.empty((n, j, 2), dtype=torch.float32, device=dev) if center is not None else None
    lib = _lib.load()
    with torch.cuda.device(dev):
        _lib.check(lib.otp_final_preds(
            _lib.dptr(hm), n, j, h, w, _lib.dptr(center, allow_none=True), _lib.dptr(scale, allow_none=True),
            idx.data_ptr(), coords.data_ptr(), preds.data_ptr() if preds is not None else None,
            maxvals.data_ptr(), _lib.stream_ptr(dev)), "otp_final_preds")
    return dict(idx=idx, coords=coords, preds=preds, maxvals=maxvals)


def _to_cuda(a, device):
    return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32)).to(device)


def get_max_preds(batch_heatmaps, device="cuda"):
    """utils/heatmap.py:143-171: returns (preds (N,J,2) float32, maxvals (N,J,1))."""
    assert batch_heatmaps.ndim == 4, 'batch_images should be 4-ndim'
    hm = _to_cuda(batch_heatmaps, device)
    r = final_preds_cuda(hm)
    idx = r["idx"].cpu().numpy().astype(np.int64)
    maxvals = r["maxvals"].cpu().numpy()
    w = batch_heatmaps.shape[3]
    preds = np.stack([idx % w, idx // w], axis=2).astype(np.float32)
    preds *= np.greater(maxvals, 0.0).astype(np.float32)
    return preds, maxvals


def get_final_preds(batch_heatmaps, center, scale, device="cuda"):
    """utils/heatmap.py:108-132: returns (preds (N,J,2) float32, maxvals (N,J,1))."""
    hm = _to_cuda(batch_heatmaps, device)
    r = final_preds_cuda(hm, _to_cuda(center, device), _to_cuda(scale, device))
    return r["preds"].cpu().numpy(), r["maxvals"].cpu().numpy()
