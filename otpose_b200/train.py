"""Training step of the temporal head (BASELINE configs[3]; reference script/Common.py:98-144, train.py:78-79).

One process per GPU.  The reference wraps the model in ``nn.DataParallel`` (scatter, replicate, gather, a
single-process all-reduce of the replicas' gradients); here every rank runs the full step on its own shard
of clips and the only exchange is ONE gradient all-reduce over NCCL (NVLink 5 / NVSwitch), bucketed and
launched from autograd hooks so that it runs under the rest of the backward pass:

* all trainable gradients live in one flat fp32 buffer (``param.grad`` are views into it), cut into buckets
  of ~``bucket_bytes`` in reverse registration order -- the order backward produces them;
* when the last gradient of a bucket has been accumulated, ``all_reduce(bucket, async_op=True)`` is issued:
  NCCL orders it behind the producing kernels and runs it on its own stream;
* ``finish()`` waits for the buckets, averages over the ranks; the global gradient norm for
  ``clip_grad_norm_`` is taken AFTER the reduction (script/Common.py:138-142), from the flat buffer.

BatchNorm in the RSB blocks uses per-rank batch statistics, like the reference's DataParallel replicas (no
SyncBN).  What runs natively and what runs through library ops in the training path: model/train_ops.py.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

__all__ = ["BucketedGradReducer", "train_step"]


class BucketedGradReducer:
    def __init__(self, params, bucket_bytes: int = 4 << 20, process_group=None):
        self.params = [p for p in params if p.requires_grad][::-1]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        total = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.buckets = []            # (start, end, [param indices])
        self._offsets = []
        o, b0, members, limit = 0, 0, [], max(1, bucket_bytes // 4)
        for i, p in enumerate(self.params):
            self._offsets.append(o)
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            members.append(i)
            o += p.numel()
            if o - b0 >= limit:
                self.buckets.append((b0, o, members))
                b0, members = o, []
        if members:
            self.buckets.append((b0, o, members))
        self._bucket_of = {}
        for bi, (_, _, mem) in enumerate(self.buckets):
            for i in mem:
                self._bucket_of[i] = bi
        self._pending = [len(m) for _, _, m in self.buckets]
        self._handles = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)
        self.enabled = True          # set False for the non-final micro-batches of a gradient accumulation
        self._hooks = [p.register_post_accumulate_grad_hook(self._make_hook(i)) for i, p in enumerate(self.params)]

    def _make_hook(self, i):
        def hook(_param):
            if not self.enabled:
                return
            bi = self._bucket_of[i]
            self._pending[bi] -= 1
            if self._pending[bi] == 0:
                self._launch(bi)
        return hook

    def _launch(self, bi):
        if self._launched[bi]:
            return
        self._launched[bi] = True
        if self.world > 1:
            s, e, _ = self.buckets[bi]
            self._handles[bi] = dist.all_reduce(self.flat[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def zero_grad(self):
        """Zero the flat buffer and re-arm the buckets (call once per optimizer step, before the first backward)."""
        self.flat.zero_()
        for i, p in enumerate(self.params):       # an optimizer's zero_grad(set_to_none=True) would drop the views
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * self._offset(i):
                p.grad = self.flat[self._offset(i):self._offset(i) + p.numel()].view_as(p)
        self._pending = [len(m) for _, _, m in self.buckets]
        self._handles = [None] * len(self.buckets)
        self._launched = [False] * len(self.buckets)

    def _offset(self, i):
        return self._offsets[i]

    def finish(self):
        """Reduce whatever has not been launched by a hook (parameters without gradient flow), wait for every
        bucket and average over the ranks."""
        for bi in range(len(self.buckets)):
            self._launch(bi)
        for h in self._handles:
            if h is not None:
                h.wait()
        if self.world > 1:
            self.flat.div_(self.world)

    def clip_grad_norm_(self, max_norm: float) -> torch.Tensor:
        """Global L2 norm of the (already reduced) gradients and in-place clipping; no host synchronisation."""
        norm = torch.linalg.vector_norm(self.flat)
        if max_norm and max_norm > 0:
            self.flat.mul_(torch.clamp(max_norm / (norm + 1e-6), max=1.0))
        return norm

    @property
    def num_buckets(self):
        return len(self.buckets)


def train_step(model, criterion, optimizer, reducer, rough_heatmaps, margin, target, target_weight,
               clip_grad_l2norm: float = 0.0, micro_batch: int = 0):
    """One optimisation step of the head on this rank's clips (script/Common.py:118-144): student loss on the
    refined heat maps against the targets and the teacher (the backbone's current-frame maps), plus the
    context-encoding term against the occlusion target, backward, gradient all-reduce, clip, optimizer step.
    ``micro_batch`` > 0 accumulates the gradient over slices of that many clips (BatchNorm statistics are then
    per slice); the all-reduce runs under the LAST slice's backward."""
    b = margin.shape[0]
    frames = rough_heatmaps.shape[0] // b
    mb = micro_batch if micro_batch and micro_batch < b else b
    reducer.zero_grad()
    per_frame = rough_heatmaps.view(frames, b, *rough_heatmaps.shape[1:])
    total = None
    starts = list(range(0, b, mb))
    for s in starts:
        e = min(b, s + mb)
        reducer.enabled = s == starts[-1]
        rough = per_frame[:, s:e].reshape(frames * (e - s), *rough_heatmaps.shape[1:])
        outs = model.forward_head(rough, margin[s:e])
        pred_s, pred_t = outs[0], rough[:e - s]
        loss = criterion(pred_s, pred_t, target[s:e], target_weight[s:e])["final_loss"]
        occlusion = (target[s:e] + outs[2]) / 2
        loss = loss + criterion(outs[4], outs[4], occlusion, target_weight[s:e])["final_loss"]
        loss = loss * ((e - s) / b)
        loss.backward()
        total = loss.detach() if total is None else total + loss.detach()
    reducer.enabled = True
    reducer.finish()
    norm = reducer.clip_grad_norm_(clip_grad_l2norm)
    optimizer.step()
    return total, norm
