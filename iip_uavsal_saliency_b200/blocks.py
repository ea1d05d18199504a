"""BasicConv2d / dwBlock (model.py:65-103) with reference-identical parameters and kernel-plan emission."""
from __future__ import annotations

import torch
import torch.nn as nn

from ._kernel_module import KernelModule
from .engine import F_RELU6, FMT_F32, FMT_Q16, FMT_SPLIT, Buf, Plan, W, fold_bn, out_size, pack_dw

__all__ = ["BasicConv2d", "dwBlock", "init_weights", "emit_stem"]

_INIT = {
    "uniform": nn.init.uniform_, "normal": nn.init.normal_, "constant": nn.init.constant_,
    "xavier_uniform": nn.init.xavier_uniform_, "xavier_normal": nn.init.xavier_normal_,
    "kaiming_uniform": nn.init.kaiming_uniform_, "kaiming_normal": nn.init.kaiming_normal_,
    "orthogonal": nn.init.orthogonal_, "sparse": nn.init.sparse_, "ones": nn.init.ones_, "zeros": nn.init.zeros_,
}
init_func = _INIT


def init_weights(model, funcname="kaiming_normal", **kwargs):
    """model.py:49-60 — conv weights by ``funcname``, BN to (1, 0), Linear to N(0, 0.01)."""
    fn = _INIT[funcname]
    for m in model.modules():
        if isinstance(m, (nn.Conv2d, nn.Conv3d)):
            fn(m.weight, **kwargs)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, (nn.BatchNorm2d, nn.BatchNorm3d)):
            nn.init.ones_(m.weight)
            nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0, 0.01)
            nn.init.zeros_(m.bias)


class BasicConv2d(nn.Sequential):
    """conv (no bias) + BatchNorm2d + ReLU6, keys ``0.weight`` / ``1.*`` (model.py:65-72)."""

    def __init__(self, in_planes, out_planes, kernel_size=3, stride=1, dilation=1, groups=1):
        padding = dilation * (kernel_size - 1) // 2
        super().__init__(
            nn.Conv2d(in_planes, out_planes, kernel_size, stride, padding, dilation=dilation, groups=groups, bias=False),
            nn.BatchNorm2d(out_planes),
            nn.ReLU6(inplace=True),
        )
        self.spec = (in_planes, out_planes, kernel_size, stride, dilation, groups)

    # ---- plan emission ----
    def folded(self):
        """(w', b') with the BatchNorm folded in, as torch tensors (reference form; the plans use ``wspec``)."""
        return fold_bn(self[0].weight, self[1])

    def wspec(self) -> W:
        """This layer's parameters for ``Plan.packed`` (BN fold + layout + hi/lo split happen in uavsal_pack_weights)."""
        return W(self[0].weight, bn=self[1], owner=self[0])

    def _emit(self, plan: Plan, x: Buf, n, h, w, out: Buf = None, res: Buf = None, tag="", out_fmt: int = FMT_SPLIT):
        cin, cout, k, stride, dil, groups = self.spec
        ws = self.wspec()
        if k == 1:
            assert stride == 1 and groups == 1
            if out is None:
                out = plan.alloc_hidden(n * h * w, cout, out_fmt)
            plan.pw(x, n * h * w, ws, None, F_RELU6, out, res=res, tag=tag)
            return out, h, w
        if groups == cin and groups == cout:
            ho, wo = out_size(h, stride), out_size(w, stride)
            out = out if out is not None else plan.alloc(n * ho * wo, cout)
            plan.dw(x, n, h, w, cout, stride, dil, ws, None, True, out, tag=tag)
            return out, ho, wo
        if groups == 1 and stride == 1 and dil == 1 and k == 3:
            out = out if out is not None else plan.alloc(n * h * w, cout)
            plan.conv3x3(x, n, h, w, cin, ws, None, F_RELU6, out, tag=tag)
            return out, h, w
        raise NotImplementedError("BasicConv2d%r has no sm_100a kernel on the UAVSal path" % (self.spec,))

    def forward(self, x):
        return _forward_single(self, x)


def emit_stem(plan: Plan, stem: BasicConv2d, x_src: torch.Tensor, kind: int, n, h, w):
    """features[0]: (normalise +) conv3x3 s2 (3->32) + BN + ReLU6 straight from the NCHW/NHWC input tensor."""
    ho, wo = out_size(h, 2), out_size(w, 2)
    out = plan.alloc_f32(n * ho * wo, 32) if plan.f32_hidden else plan.alloc(n * ho * wo, 32)   # features.1 starts with its depthwise conv
    plan.stem(x_src, kind, n, h, w, stem.wspec(), None, out, tag="features.0")
    return out, ho, wo


class dwBlock(KernelModule):
    """Inverted residual: 1x1 expand + BN + ReLU6 -> depthwise 3x3 + BN + ReLU6 -> 1x1 project + BN (+ x)
    (model.py:74-103; identical key layout to torchvision's InvertedResidual)."""

    def __init__(self, inp, oup, kernel_size=3, stride=1, expand_ratio=6, dilation=1, res_connect=None):
        super().__init__()
        assert stride in [1, 2]
        self.stride = stride
        hidden = int(round(inp * expand_ratio))
        self.use_res_connect = stride == 1 and inp == oup
        if res_connect is not None:
            self.use_res_connect = bool(res_connect and self.use_res_connect)
        seq = []
        if expand_ratio != 1:
            seq.append(BasicConv2d(inp, hidden, kernel_size=1))
        seq += [BasicConv2d(hidden, hidden, kernel_size, stride=stride, dilation=dilation, groups=hidden),
                nn.Conv2d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup)]
        self.conv = nn.Sequential(*seq)
        self.geom = (inp, oup, hidden, stride, dilation, expand_ratio != 1)

    def project_folded(self):
        conv, bn = self.conv[-2], self.conv[-1]
        wf, bf = fold_bn(conv.weight, bn)
        return wf.reshape(wf.shape[0], wf.shape[1]), bf

    def project_wspec(self) -> W:
        return W(self.conv[-2].weight, bn=self.conv[-1], owner=self.conv[-2])

    def _emit(self, plan: Plan, x: Buf, n, h, w, out: Buf = None, extra_res: Buf = None, tag=""):
        """``out`` lets the caller place the result in a concat slot.  Returns (Buf, h', w')."""
        inp, oup, hidden, stride, dil, has_expand = self.geom
        cur = x
        i = 0
        fuse_all = getattr(plan, "fuse_mbconv", True)
        if (fuse_all and has_expand and plan.engine == "tc" and stride == 1 and dil == 1 and not x.plain and x.c <= 64
                and (hidden + 63) // 64 * 64 * 3 <= hidden * 4 and oup % 8 == 0 and oup <= 64):
            # narrow stride-1 block: expand -> depthwise -> project (+ x) in ONE kernel, the 6x hidden tensor stays on the SM (mbconv.cu);
            # a hidden width that is not a multiple of 64 is zero-padded (at most a third more chunk work: 144 -> 192 for features.3)
            out = out if out is not None else plan.alloc(n * h * w, oup)
            plan.mbconv(x, n, h, w, self.conv[0].wspec(), self.conv[1].wspec(), self.project_wspec(), out,
                        res=x if self.use_res_connect else None, tag=tag + ".expand+dw+project")
            return out, h, w
        fuse = getattr(plan, "fuse_expand_dw", "auto")
        if fuse == "auto":       # measured (profiles/r01_microbench_expdw.txt): the fused kernel wins on the stride-2 high-resolution blocks
            fuse = stride == 2 and h * w >= 10000       # per-frame size: the choice must not depend on how many frames are batched
        if has_expand and plan.engine == "tc" and dil == 1 and x.c <= 32 and not x.plain and fuse:
            # few input channels: expand + depthwise in one kernel, the 6x hidden tensor never reaches HBM
            ho, wo = out_size(h, stride), out_size(w, stride)
            cur = plan.alloc(n * ho * wo, hidden)
            plan.expdw(x, n, h, w, self.conv[0].wspec(), None, stride, self.conv[1].wspec(), None, cur, tag=tag + ".expand+dw")
            if oup % 8:
                raise NotImplementedError("project conv with %d outputs is emitted by the readout path" % oup)
            out = out if out is not None else plan.alloc(n * ho * wo, oup)
            plan.pw(cur, n * ho * wo, self.project_wspec(), None, 0, out, res=x if self.use_res_connect else None, tag=tag + ".project")
            return out, ho, wo
        if has_expand:
            # the 6x hidden tensor travels as plain rows between the expand GEMM and the TMA depthwise kernels (dilation 1 only):
            # fp32, or 16-bit fixed point for the widest blocks (Plan.hidden_fmt)
            fmt = plan.hidden_fmt(hidden, dil, h * w)
            if fmt == FMT_Q16 and stride == 1 and 8 <= oup < 128:
                # dw_project with a narrow N tile is bound by its depthwise stage, where the q16 decode costs more than the bytes
                # save (1152 -> 64 @ 432 000 px: 779 vs 693 us; the expand GEMM only gains 60 us): fp32 rows
                fmt = FMT_F32
            cur, _, _ = self.conv[0]._emit(plan, cur, n, h, w, tag=tag + ".expand", out_fmt=fmt)
            i = 1
        fuse_dp = getattr(plan, "fuse_dw_project", "auto")
        if fuse_dp == "auto":    # big stride-1 blocks: the depthwise output is the project GEMM's A operand, built in shared memory
            fuse_dp = h * w >= 3600 and n * h * w >= 32768
        wide = has_expand and hidden % 128 == 0 and oup % 64 == 0 and oup <= 256          # tcgen05 pair kernel (dwproj.cu)
        narrow = (hidden, oup) == (32, 16) and not self.use_res_connect                    # features.1: fp32 FFMA kernel (dwproj32.cu)
        if fuse_dp and cur.plain and plan.engine == "tc" and stride == 1 and dil == 1 and (wide or narrow):
            out = out if out is not None else plan.alloc(n * h * w, oup)
            plan.dwproj(cur, n, h, w, self.conv[i].wspec(), None, self.project_wspec(), None, out, res=x if self.use_res_connect else None,
                        tag=tag + ".dw+project")
            return out, h, w
        cur, ho, wo = self.conv[i]._emit(plan, cur, n, h, w, tag=tag + ".dw")
        if oup % 8:
            raise NotImplementedError("project conv with %d outputs is emitted by the readout path" % oup)
        out = out if out is not None else plan.alloc(n * ho * wo, oup)
        plan.pw(cur, n * ho * wo, self.project_wspec(), None, 0, out, res=x if self.use_res_connect else None, tag=tag + ".project")
        return out, ho, wo

    def forward(self, x):
        return _forward_single(self, x)


def _forward_single(mod, x):
    """Stand-alone call of a block on an NCHW fp32 CUDA tensor (pack -> kernels -> unpack)."""
    from ._kernel_module import KernelModule, require_cuda
    require_cuda(x, type(mod).__name__)
    if isinstance(mod, KernelModule):
        return mod._forward_nchw(x)
    # BasicConv2d is an nn.Sequential (key layout); borrow the machinery through a throw-away holder
    holder = mod.__dict__.get("_holder")
    if holder is None:
        holder = _Holder(mod)
        mod.__dict__["_holder"] = holder
    return holder._forward_nchw(x)


class _Holder(KernelModule):
    def __init__(self, inner):
        super().__init__()
        self.__dict__["_inner"] = inner

    def parameters(self, recurse=True):
        return self.__dict__["_inner"].parameters(recurse)

    def buffers(self, recurse=True):
        return self.__dict__["_inner"].buffers(recurse)

    def _emit(self, plan, x, n, h, w):
        inner = self.__dict__["_inner"]
        if inner.spec[2] == 3 and inner.spec[5] == 1 and inner.spec[0] == 3:
            raise NotImplementedError("the 3-channel stem runs from the NCHW input (see ReMobileNetV2)")
        return inner._emit(plan, x, n, h, w)
