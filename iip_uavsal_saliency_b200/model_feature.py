"""Drop-in for the reference's model_feature.py: ReMobileNetV2 (the hot path, model_feature.py:49-69) and the alternative
backbones ReResNet (:72-103) / ReVGG (:106-128) that ``uavsal_srfnet_aspp(cnn_type=...)`` can be built on (model.py:14-33).

The MobileNetV2 trunk is built here from the same two block types the rest of the model uses, so the
state-dict keys equal torchvision's ``features.{i}...`` layout without importing torchvision:
``features.0`` conv-BN-ReLU6 stem, ``features.1`` the t=1 block, ``features.2..17`` t=6 inverted residuals,
``features.18`` the 320->1280 head that the reference keeps in the state dict but never executes
(model_feature.py:68).  The ResNets / VGG-16 are rebuilt the same way with torchvision's attribute names
(``conv1 / bn1 / layerN.M.convK / bnK / downsample.{0,1}``; ``features.N`` for VGG), without importing torchvision; their
classification heads (avgpool / fc / classifier), which the reference discards, are not created.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ._kernel_module import KernelModule
from .blocks import BasicConv2d, dwBlock, emit_stem
from .engine import F_RELU, W, out_size

__all__ = ["ReMobileNetV2", "ReResNet", "ReVGG", "feature_loader"]

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


# ---------------------------------------------------------------------------------------------------
# ResNet (torchvision resnet.py) behind ReResNet (model_feature.py:72-103)
# ---------------------------------------------------------------------------------------------------
_RESNET_CFG = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet34": ("basic", (3, 4, 6, 3)), "resnet50": ("bottleneck", (3, 4, 6, 3)),
               "resnet101": ("bottleneck", (3, 4, 23, 3)), "resnet152": ("bottleneck", (3, 8, 36, 3))}
# torchvision.models.resnet.__all__ / vgg.__all__ model names (model_feature.py:8-10: anything else is a ValueError)
_RESNET_NAMES = ("resnet18", "resnet34", "resnet50", "resnet101", "resnet152", "resnext50_32x4d", "resnext101_32x8d", "resnext101_64x4d",
                 "wide_resnet50_2", "wide_resnet101_2", "ResNet")
_VGG_NAMES = ("VGG", "vgg11", "vgg11_bn", "vgg13", "vgg13_bn", "vgg16", "vgg16_bn", "vgg19", "vgg19_bn")


def _pool(plan, x, n, h, w, k, stride, pad, tag):
    ho, wo = (h + 2 * pad - k) // stride + 1, (w + 2 * pad - k) // stride + 1
    out = plan.alloc(n * ho * wo, x.c)
    plan.maxpool(x, n, h, w, x.c, k, stride, pad, out, tag=tag)
    return out, ho, wo


class _ResBlock(nn.Module):
    """torchvision BasicBlock (two 3x3 convs) / Bottleneck (1x1, 3x3 with the stride, 1x1 x4), attribute names as torchvision."""

    def __init__(self, kind, inplanes, planes, stride):
        super().__init__()
        self.kind, self.stride = kind, stride
        exp = 1 if kind == "basic" else 4
        if kind == "basic":
            self.conv1 = nn.Conv2d(inplanes, planes, 3, stride, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.relu = nn.ReLU(inplace=True)
            self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
        else:
            self.conv1 = nn.Conv2d(inplanes, planes, 1, bias=False)
            self.bn1 = nn.BatchNorm2d(planes)
            self.conv2 = nn.Conv2d(planes, planes, 3, stride, 1, bias=False)
            self.bn2 = nn.BatchNorm2d(planes)
            self.conv3 = nn.Conv2d(planes, planes * 4, 1, bias=False)
            self.bn3 = nn.BatchNorm2d(planes * 4)
            self.relu = nn.ReLU(inplace=True)
        self.downsample = None
        if stride != 1 or inplanes != planes * exp:
            self.downsample = nn.Sequential(nn.Conv2d(inplanes, planes * exp, 1, stride, bias=False), nn.BatchNorm2d(planes * exp))
        self.out_channels = planes * exp

    def _conv3(self, plan, x, n, h, w, conv, bn, stride, relu, tag):
        """3x3 conv (+BN, +ReLU).  Stride 2 = the stride-1 conv subsampled (output (y, x) of a pad-1 stride-2 conv is the stride-1
        output at (2y, 2x)): three or four layers per ResNet, not worth a strided implicit-GEMM loader."""
        y = plan.alloc(n * h * w, conv.out_channels)
        plan.conv3x3(x, n, h, w, conv.in_channels, W(conv.weight, bn=bn, owner=conv), None, F_RELU if relu else 0, y, tag=tag)
        if stride == 2:
            return _pool(plan, y, n, h, w, 1, 2, 0, tag + ".s2")
        return y, h, w

    def _emit(self, plan, x, n, h, w, tag):
        s = self.stride
        if self.kind == "basic":
            y, ho, wo = self._conv3(plan, x, n, h, w, self.conv1, self.bn1, s, True, tag + ".conv1")
            y, _, _ = self._conv3(plan, y, n, ho, wo, self.conv2, self.bn2, 1, False, tag + ".conv2")
        else:
            y = plan.alloc(n * h * w, self.conv1.out_channels)
            plan.pw(x, n * h * w, W(self.conv1.weight, bn=self.bn1, owner=self.conv1), None, F_RELU, y, tag=tag + ".conv1")
            y, ho, wo = self._conv3(plan, y, n, h, w, self.conv2, self.bn2, s, True, tag + ".conv2")
            y3 = plan.alloc(n * ho * wo, self.out_channels)
            plan.pw(y, n * ho * wo, W(self.conv3.weight, bn=self.bn3, owner=self.conv3), None, 0, y3, tag=tag + ".conv3")
            y = y3
        idn = x
        if self.downsample is not None:
            xs = x
            if s == 2:
                xs, _, _ = _pool(plan, x, n, h, w, 1, 2, 0, tag + ".ds.s2")
            idn = plan.alloc(n * ho * wo, self.out_channels)
            plan.pw(xs, n * ho * wo, W(self.downsample[0].weight, bn=self.downsample[1], owner=self.downsample[0]), None, 0, idn, tag=tag + ".downsample")
        out = plan.alloc(n * ho * wo, self.out_channels)
        plan.add_act(y, idn, n * ho * wo, self.out_channels, F_RELU, out, tag=tag + ".add+relu")
        return out, ho, wo


class _LevelsModule(KernelModule):
    """forward(x) -> the five pyramid levels as NCHW fp32 tensors, through a cached kernel plan."""

    def forward(self, x):
        from ._kernel_module import require_cuda
        require_cuda(x, type(self).__name__)
        n, c, h, w = x.shape
        if c != 3:
            raise RuntimeError("%s expects 3 input channels, got %d" % (type(self).__name__, c))

        def build(plan):
            xin = plan.tensor((n, 3, h, w))
            outs = []
            for buf, hh, ww in self._emit_levels(plan, xin, 0, n, h, w):
                t = plan.tensor((n, buf.c, hh, ww))
                plan.unpack_nchw(buf, n, buf.c, hh, ww, t)
                outs.append(t)
            plan.named.update(x_in=xin, outs=outs)

        plan = self._cached_plan((x.device, "levels", n, h, w), build)
        plan.named["x_in"].copy_(x)
        plan.launch()
        return tuple(t.clone() for t in plan.named["outs"])


class ReResNet(_LevelsModule):
    """model_feature.py:72-103: conv1 / bn1 / relu / maxpool / layer1..4 of a torchvision ResNet; forward returns
    (x0 after the max pool, layer1, layer2, layer3, layer4)."""

    def __init__(self, name="resnet50"):
        super().__init__()
        if name not in _RESNET_NAMES:
            raise ValueError                         # model_feature.py:75-76
        if name.lower() not in _RESNET_CFG:
            raise NotImplementedError                # model_feature.py:77-78 (resnext / wide variants are not in feature_loader)
        kind, blocks = _RESNET_CFG[name.lower()]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        self.relu = nn.ReLU(inplace=True)
        self.maxpool = nn.MaxPool2d(3, 2, 1)
        inpl = 64
        for li, nb in enumerate(blocks):
            planes = 64 << li
            layer = []
            for bi in range(nb):
                blk = _ResBlock(kind, inpl, planes, 2 if (bi == 0 and li > 0) else 1)
                inpl = blk.out_channels
                layer.append(blk)
            setattr(self, "layer%d" % (li + 1), nn.Sequential(*layer))
        for m in self.modules():                     # torchvision's init
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.ones_(m.weight)
                nn.init.zeros_(m.bias)

    def _emit_levels(self, plan, x_src, kind, n, h, w):
        ho, wo = (h + 6 - 7) // 2 + 1, (w + 6 - 7) // 2 + 1
        cur = plan.alloc(n * ho * wo, 64)
        plan.conv_first(x_src, kind, n, h, w, W(self.conv1.weight, bn=self.bn1, owner=self.conv1), 2, F_RELU, cur, tag="conv1")
        cur, ch, cw = _pool(plan, cur, n, ho, wo, 3, 2, 1, "maxpool")
        levels = [(cur, ch, cw)]
        for li in range(4):
            for bi, blk in enumerate(getattr(self, "layer%d" % (li + 1))):
                cur, ch, cw = blk._emit(plan, cur, n, ch, cw, "layer%d.%d" % (li + 1, bi))
            levels.append((cur, ch, cw))
        return levels


class ReVGG(_LevelsModule):
    """model_feature.py:106-128 over torchvision's vgg16.features (conv3x3 + bias + ReLU, 2x2 max pools).  The reference's split
    points sit one past every max pool (it enumerates ``features.modules()``, whose first element is the Sequential itself), so
    each of the five levels ends with its pooling layer."""
    _CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")

    def __init__(self, name="vgg16"):
        super().__init__()
        if name not in _VGG_NAMES:
            raise ValueError                         # model_feature.py:110-111
        if name.lower() != "vgg16":
            raise NotImplementedError                # model_feature.py:112-113 (only vgg16 is in feature_loader)
        layers, cin = [], 3
        for v in self._CFG:
            if v == "M":
                layers.append(nn.MaxPool2d(2, 2))
            else:
                layers += [nn.Conv2d(cin, v, 3, padding=1), nn.ReLU(inplace=True)]
                cin = v
        self.features = nn.Sequential(*layers)
        for m in self.features:
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                nn.init.zeros_(m.bias)

    def _emit_levels(self, plan, x_src, kind, n, h, w):
        levels, cur, ch, cw = [], None, h, w
        for i, m in enumerate(self.features):
            if isinstance(m, nn.Conv2d):
                ws = W(m.weight, bias=m.bias, owner=m)
                out = plan.alloc(n * ch * cw, m.out_channels)
                if cur is None:
                    plan.conv_first(x_src, kind, n, h, w, ws, 1, F_RELU, out, tag="features.0")
                else:
                    plan.conv3x3(cur, n, ch, cw, m.in_channels, ws, None, F_RELU, out, tag="features.%d" % i)
                cur = out
            elif isinstance(m, nn.MaxPool2d):
                cur, ch, cw = _pool(plan, cur, n, ch, cw, 2, 2, 0, "features.%d" % i)
                levels.append((cur, ch, cw))
        return levels


for _n in _RESNET_CFG:
    feature_loader[_n] = None                       # built by ReResNet itself (kept for the reference's `name in feature_loader` checks)
feature_loader["vgg16"] = None
