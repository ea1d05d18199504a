"""Drop-in for the reference's model_feature.py (hot path only: ReMobileNetV2, model_feature.py:49-69).

The MobileNetV2 trunk is built here from the same two block types the rest of the model uses, so the
state-dict keys equal torchvision's ``features.{i}...`` layout without importing torchvision:
``features.0`` conv-BN-ReLU6 stem, ``features.1`` the t=1 block, ``features.2..17`` t=6 inverted residuals,
``features.18`` the 320->1280 head that the reference keeps in the state dict but never executes
(model_feature.py:68).  ReResNet / ReVGG are out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._kernel_module import KernelModule
from .blocks import BasicConv2d, dwBlock, emit_stem

__all__ = ["ReMobileNetV2", "feature_loader"]

# (expand t, channels c, repeats n, first stride s) — MobileNetV2 table 2
_SETTING = ((1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1))
# feature taps returned by forward: indices of the LAST block of each slice [0:2],[2:4],[4:7],[7:14],[14:18]
_TAPS = (1, 3, 6, 13, 17)


def _mobilenet_v2_features() -> nn.Sequential:
    layers = [BasicConv2d(3, 32, 3, stride=2)]
    inp = 32
    for t, c, n, s in _SETTING:
        for i in range(n):
            layers.append(dwBlock(inp, c, 3, stride=s if i == 0 else 1, expand_ratio=t))
            inp = c
    layers.append(BasicConv2d(inp, 1280, 1))
    feats = nn.Sequential(*layers)
    for m in feats.modules():                       # torchvision's own init (kaiming fan_out, BN 1/0)
        if isinstance(m, nn.Conv2d):
            nn.init.kaiming_normal_(m.weight, mode="fan_out")
        elif isinstance(m, nn.BatchNorm2d):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
    return feats


feature_loader = {"mobilenet_v2": _mobilenet_v2_features}


class ReMobileNetV2(KernelModule):
    def __init__(self, name="mobilenet_v2"):
        super().__init__()
        if name not in ("mobilenet_v2", "MobileNetV2", "MobileNet_V2_Weights"):
            raise ValueError                         # model_feature.py:54-55
        if name.lower() not in feature_loader:
            raise NotImplementedError                # model_feature.py:56-57
        self.features = feature_loader[name.lower()]()

    # plan emission: returns the five pyramid levels as (Buf, h, w)
    def _emit_levels(self, plan, x_src, kind, n, h, w):
        cur, ch, cw = emit_stem(plan, self.features[0], x_src, kind, n, h, w)
        levels = []
        for i in range(1, 18):
            cur, ch, cw = self.features[i]._emit(plan, cur, n, ch, cw, tag="features.%d" % i)
            if i in _TAPS:
                levels.append((cur, ch, cw))
        return levels

    def forward(self, x):
        from ._kernel_module import require_cuda
        require_cuda(x, "ReMobileNetV2")
        n, c, h, w = x.shape
        if c != 3:
            raise RuntimeError("ReMobileNetV2 expects 3 input channels, got %d" % c)

        def build(plan):
            xin = plan.tensor((n, 3, h, w))
            outs = []
            for buf, hh, ww in self._emit_levels(plan, xin, 0, n, h, w):
                t = plan.tensor((n, buf.c, hh, ww))
                plan.unpack_nchw(buf, n, buf.c, hh, ww, t)
                outs.append(t)
            plan.named.update(x_in=xin, outs=outs)

        plan = self._cached_plan((x.device, "mbv2", n, h, w), build)
        plan.named["x_in"].copy_(x)
        plan.launch()
        return tuple(t.clone() for t in plan.named["outs"])
