"""Drop-in ``HRNet`` backbone (reference model/HRNet.py:14-152 constructor / forward, :391-595 modules).

SURVEY.md section 8f rank 3 / BASELINE configs[1]: the backbone is cuDNN work, not part of the hand-written
hot path -- it is here so that the FULL inference (HRNet-W48 + temporal head) can be run and timed without
the reference tree, and so that the backbone -> head hand-off of section 8f rank 1 has a real producer.
Same constructor (``HRNet(cfg)`` with ``cfg.MODEL.EXTRA.STAGE{2,3,4}``), same state-dict keys and shapes as
the reference (checked against a manifest dumped from the reference's own module,
``tests/golden/state_dict_manifest.json["hrnet"]``) so its checkpoints load with ``load_state_dict``.

Inference path (``.eval()``): every Conv2d + BatchNorm2d pair is folded into one convolution with bias
(``fold()``), tensors run channels-last in the module's dtype (bf16 on B200), the ReLU / residual adds are
the only elementwise launches left, and ``features(x)`` returns the last high-resolution map ``y_list[0]``
(48 channels) for ``OTPose.forward_from_features``, which applies ``final_layer`` inside the head's first
kernel.  ``forward(x)`` returns the rough heat maps exactly like the reference.

Structure (Sun et al., HRNet-W48; layer names follow the reference state dict):
  stem conv1/bn1, conv2/bn2 (stride 2 each) -> layer1: 4 Bottlenecks (64 -> 256)
  transition1 -> stage2 (1 module, 2 branches 48/96) -> transition2 -> stage3 (4 modules, 3 branches
  48/96/192) -> transition3 -> stage4 (3 modules, 4 branches 48/96/192/384; last module fuses to branch 0
  only) -> final_layer 1x1 (48 -> joints).
"""
from __future__ import annotations

import torch
import torch.nn.functional as F
from torch import nn

BN_MOMENTUM = 0.1
__all__ = ["HRNet", "HighResolutionModule", "hrnet_w48_cfg"]


def _conv_bn(cin, cout, k, stride, relu):
    layers = [nn.Conv2d(cin, cout, k, stride, k // 2, bias=False), nn.BatchNorm2d(cout, momentum=BN_MOMENTUM)]
    if relu:
        layers.append(nn.ReLU(inplace=True))
    return nn.Sequential(*layers)


class Interpolate(nn.Module):
    """Nearest-neighbour upsampling by an integer factor (parameter-free member of the fuse layers)."""

    def __init__(self, scale_factor, mode="nearest"):
        super().__init__()
        self.scale_factor, self.mode = scale_factor, mode

    def forward(self, x):
        return F.interpolate(x, scale_factor=self.scale_factor, mode=self.mode)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, groups=groups, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.act_fun = nn.ReLU(inplace=True)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        res = x if self.downsample is None else self.downsample(x)
        y = self.act_fun(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.act_fun(y + res)


class Bottleneck(nn.Module):
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes, momentum=BN_MOMENTUM)
        self.conv3 = nn.Conv2d(planes, planes * self.expansion, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * self.expansion, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        res = x if self.downsample is None else self.downsample(x)
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + res)


blocks_dict = {"BASIC": BasicBlock, "BOTTLENECK": Bottleneck}


def _branch(block, cin, planes, n_blocks):
    ds = None
    if cin != planes * block.expansion:
        ds = _conv_bn(cin, planes * block.expansion, 1, 1, relu=False)
    layers = [block(cin, planes, 1, ds)]
    layers += [block(planes * block.expansion, planes) for _ in range(1, n_blocks)]
    return nn.Sequential(*layers)


class HighResolutionModule(nn.Module):
    """Parallel residual branches followed by the all-to-all multi-resolution fusion (sum)."""

    def __init__(self, num_branches, blocks, num_blocks, num_inchannels, num_channels, fuse_method,
                 multi_scale_output=True):
        super().__init__()
        if not (num_branches == len(num_blocks) == len(num_channels) == len(num_inchannels)):
            raise ValueError("NUM_BRANCHES disagrees with NUM_BLOCKS / NUM_CHANNELS / input channels")
        self.num_branches, self.fuse_method, self.multi_scale_output = num_branches, fuse_method, multi_scale_output
        self.branches = nn.ModuleList(_branch(blocks, num_inchannels[i], num_channels[i], num_blocks[i])
                                      for i in range(num_branches))
        self.num_inchannels = [c * blocks.expansion for c in num_channels]
        self.fuse_layers = self._fuse_layers()
        self.relu = nn.ReLU(True)

    def get_num_inchannels(self):
        return self.num_inchannels

    def _fuse_layers(self):
        if self.num_branches == 1:
            return None
        ch = self.num_inchannels
        rows = []
        for i in range(self.num_branches if self.multi_scale_output else 1):
            row = []
            for j in range(self.num_branches):
                if j == i:
                    row.append(None)
                elif j > i:      # lower resolution -> 1x1 conv, then nearest upsampling
                    row.append(nn.Sequential(nn.Conv2d(ch[j], ch[i], 1, 1, 0, bias=False), nn.BatchNorm2d(ch[i]),
                                             Interpolate(scale_factor=2 ** (j - i), mode="nearest")))
                else:            # higher resolution -> chain of stride-2 3x3 convs, the last one changes width
                    steps = [_conv_bn(ch[j], ch[i] if k == i - j - 1 else ch[j], 3, 2, relu=k != i - j - 1)
                             for k in range(i - j)]
                    row.append(nn.Sequential(*steps))
            rows.append(nn.ModuleList(row))
        return nn.ModuleList(rows)

    def forward(self, x):
        x = [br(xi) for br, xi in zip(self.branches, x)]
        if self.num_branches == 1:
            return x
        out = []
        for i, row in enumerate(self.fuse_layers):
            y = x[0] if i == 0 else row[0](x[0])
            for j in range(1, self.num_branches):
                y = y + (x[j] if j == i else row[j](x[j]))
            out.append(self.relu(y))
        return out


class _Cfg(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e


def hrnet_w48_cfg(num_joints=17):
    """MODEL section of configs/Base_PoseTrack17.yaml:46-88 (HRNet-W48)."""
    def stage(modules, chans):
        return _Cfg(NUM_MODULES=modules, NUM_BRANCHES=len(chans), BLOCK="BASIC", NUM_BLOCKS=[4] * len(chans),
                    NUM_CHANNELS=list(chans), FUSE_METHOD="SUM")
    return _Cfg(MODEL=_Cfg(PRETRAINED="", NUM_JOINTS=num_joints, FREEZE_HRNET_WEIGHTS=True,
                           EXTRA=_Cfg(PRETRAINED_LAYERS=["*"], FINAL_CONV_KERNEL=1, STAGE2=stage(1, (48, 96)),
                                      STAGE3=stage(4, (48, 96, 192)), STAGE4=stage(3, (48, 96, 192, 384)))))


class HRNet(nn.Module):
    def __init__(self, cfg, **kwargs):
        super().__init__()
        extra = cfg["MODEL"]["EXTRA"]
        self.pretrained = cfg["MODEL"]["PRETRAINED"]
        self.freeze_hrnet_weight = cfg["MODEL"]["FREEZE_HRNET_WEIGHTS"]
        self.conv1 = nn.Conv2d(3, 64, 3, 2, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.conv2 = nn.Conv2d(64, 64, 3, 2, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(64, momentum=BN_MOMENTUM)
        self.relu = nn.ReLU(inplace=True)
        self.layer1 = _branch(Bottleneck, 64, 64, 4)
        pre = [256]
        for s in (2, 3, 4):
            sc = extra[f"STAGE{s}"]
            block = blocks_dict[sc["BLOCK"]]
            chans = [c * block.expansion for c in sc["NUM_CHANNELS"]]
            setattr(self, f"stage{s}_cfg", sc)
            setattr(self, f"transition{s - 1}", self._transition(pre, chans))
            mods = []
            cin = chans
            for m in range(sc["NUM_MODULES"]):
                multi = not (s == 4 and m == sc["NUM_MODULES"] - 1)      # the very last module fuses to branch 0 only
                mods.append(HighResolutionModule(sc["NUM_BRANCHES"], block, sc["NUM_BLOCKS"], cin, sc["NUM_CHANNELS"],
                                                 sc["FUSE_METHOD"], multi))
                cin = mods[-1].get_num_inchannels()
            setattr(self, f"stage{s}", nn.Sequential(*mods))
            pre = cin
        self.pre_stage_channels = pre
        k = extra["FINAL_CONV_KERNEL"]
        self.final_layer = nn.Conv2d(pre[0], cfg["MODEL"]["NUM_JOINTS"], k, 1, 1 if k == 3 else 0)
        self._folded = None

    @classmethod
    def get_net(cls, cfg, **kwargs):
        return cls(cfg, **kwargs)

    @staticmethod
    def _transition(pre, cur):
        layers = []
        for i, c in enumerate(cur):
            if i < len(pre):
                layers.append(_conv_bn(pre[i], c, 3, 1, relu=True) if c != pre[i] else None)
            else:            # new, lower-resolution branch: stride-2 convs from the last previous branch
                n = i + 1 - len(pre)
                layers.append(nn.Sequential(*[_conv_bn(pre[-1], c if j == n - 1 else pre[-1], 3, 2, relu=True)
                                              for j in range(n)]))
        return nn.ModuleList(layers)

    def freeze_weight(self):
        for p in self.parameters():
            p.requires_grad = False

    def init_weights(self, *args, **kwargs):
        import os.path as osp
        if self.pretrained and osp.isfile(self.pretrained):
            sd = torch.load(self.pretrained, map_location="cpu")
            self.load_state_dict(sd.get("state_dict", sd), strict=False)

    # ------------------------------------------------------------------ forward
    def _stages(self, x):
        x = self.relu(self.bn1(self.conv1(x)))
        x = self.relu(self.bn2(self.conv2(x)))
        x = self.layer1(x)
        ys = [x]
        for s in (2, 3, 4):
            # model/HRNet.py:127-146: a transition layer reads the single layer1 map (stage 2) or the LAST
            # (lowest-resolution) map of the previous stage; branches without one pass through
            tr = getattr(self, f"transition{s - 1}")
            xs = [(ys[i] if s > 2 else x) if t is None else t(ys[-1]) for i, t in enumerate(tr)]
            ys = getattr(self, f"stage{s}")(xs)
        return ys

    def features(self, x):
        """``y_list[0]`` of model/HRNet.py:149 -- the (N, 48, H/4, W/4) map that ``final_layer`` reads."""
        return self._stages(x)[0]

    def forward(self, x, **kwargs):
        return self.final_layer(self.features(x))

    # ------------------------------------------------------------------ inference packing
    @torch.no_grad()
    def fold(self, dtype=torch.bfloat16, memory_format=torch.channels_last):
        """Return an eval-only copy in which every Conv2d + BatchNorm2d pair is ONE convolution with bias
        (W' = W * g / sqrt(var + eps), b' = beta - mean * g / sqrt(var + eps)), cast to ``dtype`` and laid out
        ``memory_format`` -- the form cuDNN runs fastest on B200.  The copy no longer has the reference's
        state-dict layout; keep ``self`` for checkpoints."""
        import copy
        m = copy.deepcopy(self).eval()

        def fold_pair(conv, bn):
            g = bn.weight / torch.sqrt(bn.running_var + bn.eps)
            conv.weight.data = conv.weight.data * g.view(-1, 1, 1, 1)
            b = bn.bias - bn.running_mean * g
            if conv.bias is not None:
                b = b + conv.bias.data * g
            conv.bias = nn.Parameter(b)

        def walk(mod):
            names = list(mod._modules.keys())
            for a, b in zip(names, names[1:]):
                ca, cb = mod._modules[a], mod._modules[b]
                if isinstance(ca, nn.Conv2d) and isinstance(cb, nn.BatchNorm2d):
                    fold_pair(ca, cb)
                    mod._modules[b] = nn.Identity()
            for child in mod._modules.values():
                if child is not None:
                    walk(child)

        walk(m)
        return m.to(dtype=dtype, memory_format=memory_format)
