"""GPU: the training path of the head (BASELINE configs[3]).

Parity protocol: dropout / drop-path are stochastic and BatchNorm batch statistics change the function, so
the gradient check runs the TRAINING code path with dropout probabilities 0 and the BatchNorm layers in eval
mode -- then it must reproduce the oracle's forward (the reference semantics) and the gradients of float64
autograd through the oracle.  Train-mode specifics (drop-path scaling, BN batch statistics, seeded dropout
replay) are checked as properties.
"""
import numpy as np
import pytest
import torch
from torch import nn

from oracle import head_oracle as ho
from otpose_b200.model import OTPose, default_cfg
from otpose_b200.model.loss import ST_OHKW_MSELoss
from otpose_b200.train import BucketedGradReducer, train_step
from otpose_b200.utils import synthetic as syn

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def full_fp32_library_convs():
    """The library convolutions of the training path default to TF32 in cuDNN (1e-3-level rounding, amplified
    by the learned-offset sampling); parity against float64 autograd is checked with TF32 off."""
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def build(h, w, seed=2024, dropout=True):
    model = OTPose(default_cfg((h, w)))
    sd = syn.fill_state_dict({k: v.shape for k, v in model.state_dict().items()}, seed=seed)
    model.load_state_dict(sd)
    model = model.cuda().train()
    if not dropout:
        for m in model.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
            if hasattr(m, "drop_prob"):
                m.drop_prob = 0.0
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
    return model, sd


def test_training_path_matches_oracle_forward_and_float64_gradients():
    b, h, w = 2, 16, 12
    model, sd = build(h, w, dropout=False)
    rough = syn.synth_rough_heatmaps(b, 17, h, w)
    margin = syn.synth_margin(b)
    outs = model.forward_head(rough.cuda(), margin.cuda())
    sd64 = {k: (v.double().requires_grad_(v.is_floating_point() and "running" not in k and "pos_embd" not in k)
                if v.is_floating_point() else v) for k, v in sd.items()}
    ref = ho.head_forward(sd64, rough.double(), margin)
    names = ("output_heatmaps", "rough", "intersection", "prev_b", "context_encoding", "squeezed", "total_b")
    for n, o, r in zip(names, outs, ref):
        if n == "rough":
            continue
        err = (o.detach().cpu().double() - r.detach()).abs().max() / r.detach().abs().max()
        assert err < 1e-4, (n, float(err))
    # scalar objective touching both differentiable outputs
    g = torch.Generator().manual_seed(3)
    c0, c4 = torch.randn(outs[0].shape, generator=g), torch.randn(outs[4].shape, generator=g)
    (outs[0] * c0.cuda()).sum().add((outs[4] * c4.cuda()).sum()).backward()
    (ref[0] * c0.double()).sum().add((ref[4] * c4.double()).sum()).backward()
    checked = 0
    for name, p in model.named_parameters():
        r = sd64[name].grad
        if r is None:
            assert p.grad is None or float(p.grad.abs().max()) == 0.0, name
            continue
        scale = float(r.abs().max())
        if scale == 0.0:
            continue
        err = float((p.grad.cpu().double() - r).abs().max()) / scale
        assert err < 1e-2, (name, err)      # fp32 vs float64 through ~60 layers and the learned-offset sampling
        checked += 1
    assert checked >= 800         # every trainable tensor of the head (818 of the 968 state-dict entries are parameters)


def test_train_mode_semantics():
    b, h, w = 2, 16, 12
    model, _ = build(h, w)
    rough, margin = syn.synth_rough_heatmaps(b, 17, h, w).cuda(), syn.synth_margin(b).cuda()
    torch.cuda.manual_seed(11)
    a = model.forward_head(rough, margin)[0]
    torch.cuda.manual_seed(11)
    a2 = model.forward_head(rough, margin)[0]
    torch.cuda.manual_seed(12)
    c = model.forward_head(rough, margin)[0]
    assert torch.equal(a, a2)                       # Philox replay: same seed, same dropout / drop-path masks
    assert not torch.equal(a, c)                    # and they are really active
    bn = model.def_fuse.layers[0].conv_bn_relu1.bn
    assert int(bn.num_batches_tracked) == 3         # batch statistics: running stats updated by each call
    model.eval()
    e = model.forward_head(rough, margin)[0]        # back on the fused CUDA kernels
    assert torch.isfinite(e).all()


def test_train_step_runs_and_descends():
    """fwd + bwd + (single-rank) gradient reduction + global-norm clip + AdamW on 4 clips, micro-batched."""
    b, h, w = 4, 16, 12
    model, _ = build(h, w, dropout=False)
    params = [p for p in model.parameters() if p.requires_grad]
    red = BucketedGradReducer(params, bucket_bytes=1 << 20)
    assert red.num_buckets >= 4
    opt = torch.optim.AdamW(params, lr=1e-3)
    crit = ST_OHKW_MSELoss(use_target_weight=True)
    rough, margin = syn.synth_rough_heatmaps(b, 17, h, w).cuda(), syn.synth_margin(b).cuda()
    target = syn.synth_rough_heatmaps(b, 17, h, w, frames=1, seed=99).cuda()
    tw = torch.ones(b, 17, 1, device="cuda")
    losses = [float(train_step(model, crit, opt, red, rough, margin, target, tw, clip_grad_l2norm=1.0,
                               micro_batch=2)[0]) for _ in range(6)]
    assert all(np.isfinite(losses))
    assert losses[-1] < losses[0]
