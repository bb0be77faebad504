"""Drop-in ``ST_OHKW_MSELoss`` (reference model/loss.py:5-92) that stays on the device.

The reference loops over the joints in Python and decides per joint with ``if torch.max(heatmap_gt) == 1``
-- a device -> host synchronisation for each of the 17 joints of every step -- whether the student is also
pulled towards the teacher heat map.  Here the 17 decisions are one boolean vector on the device and the
whole loss is a handful of batched tensor ops: same value, no synchronisation, differentiable.

    per joint j:  w = target_weight[:, j]                                     (B, 1)
                  e_gt = (s_j w - gt_j w)^2 ,  e_t = (s_j w - t_j w)^2        (B, HW)
                  has_peak_j = max(gt_j) == 1
                  loss_s_j = 0.5 e_gt                     if has_peak_j else 0.5 (e_gt + e_t)
                  mse_s   += mean(e_gt)                   (+ mean(e_t) if not has_peak_j)
    ohkm = mean_b( sum(top-8_j mean_hw loss_s_j) / 8 ),  final_loss = ohkm + mse_s
"""
from __future__ import annotations

import torch
from torch import nn


class ST_OHKW_MSELoss(nn.Module):
    def __init__(self, use_target_weight, topk=8):
        super().__init__()
        if not use_target_weight:
            raise NotImplementedError("the reference's un-weighted branch builds no student loss (loss.py:73-76, 81-82)")
        self.use_target_weight = use_target_weight
        self.topk = topk

    def forward(self, output_s, output_t, target, target_weight, effective_num_joints: int = None):
        b, j = output_t.shape[0], output_t.shape[1]
        if effective_num_joints is None:
            effective_num_joints = j
        s = output_s.reshape(b, j, -1)
        t = output_t.reshape(b, j, -1)
        gt = target.reshape(b, j, -1)
        w = target_weight.reshape(b, j, 1)
        sw = s * w
        e_gt = (sw - gt * w) ** 2                                   # (B, J, HW)
        e_t = (sw - t * w) ** 2
        no_peak = (gt.amax(dim=(0, 2)) != 1).to(s.dtype)            # (J,): 1 where the teacher term applies
        loss_s = 0.5 * (e_gt + e_t * no_peak[None, :, None])        # (B, J, HW)
        mse_s = e_gt.mean(dim=(0, 2)).sum() + (e_t.mean(dim=(0, 2)) * no_peak).sum()
        per_joint = loss_s.mean(dim=2)                              # (B, J)
        ohkm = per_joint.topk(self.topk, dim=1, sorted=False).values.sum(dim=1).div(self.topk).mean()
        return {"ohkm_loss_s": ohkm, "mse_loss_s": mse_s / effective_num_joints, "final_loss": ohkm + mse_s}
