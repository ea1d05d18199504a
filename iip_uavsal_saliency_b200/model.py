"""Drop-in for the reference's model.py hot path (model.py:1-375): BasicConv2d, dwBlock, uavsal_srfnet_aspp,
spConv, teConv_sub, STBlock, UAVSal — same constructors, forward signatures and 685-key state dict, so
published weights load with ``strict=True``.  ``forward`` does not run PyTorch ops: it replays a plan of
hand-written sm_100a kernels (see engine.py / csrc/) built once per input shape.

Out of scope: the ablation zoo of model.py:376-1077.
"""
from __future__ import annotations

import os
from typing import List, Optional

import numpy as np
import torch
import torch.nn as nn

from ._kernel_module import KernelModule, require_cuda
from .blocks import BasicConv2d, dwBlock, init_func, init_weights
from .engine import F_RELU6, Buf, Plan, out_size
from .model_convlstm import *            # noqa: F401,F403  (the reference re-exports these, model.py:11)
from .model_convlstm import ConvLSTM, ConvTWA, emit_twa
from .model_feature import ReMobileNetV2, ReResNet, ReVGG

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")     # model.py:8

feature_loader = {"vgg16": ReVGG, "resnet18": ReResNet, "resnet34": ReResNet, "resnet50": ReResNet, "resnet101": ReResNet,
                  "resnet152": ReResNet, "mobilenet_v2": ReMobileNetV2}      # model.py:14-22
feature_inplanes = {"vgg16": [128, 256, 512, 512], "resnet18": [64, 128, 256, 512], "resnet34": [64, 128, 256, 512],
                    "resnet50": [256, 512, 1024, 2048], "resnet101": [256, 512, 1024, 2048], "resnet152": [256, 512, 1024, 2048],
                    "mobilenet_v2": [24, 32, 96, 320]}                      # model.py:25-33


class uavsal_srfnet_aspp(KernelModule):
    """SRF-Net: MobileNetV2 pyramid + ASPP on c5 + laterals + 3x3 fuse (model.py:110-158)."""

    def __init__(self, cnn_type="mobilenet_v2", planes=[64, 64, 128, 256], last_channel=256):
        super().__init__()
        if last_channel == 128:
            planes = [32, 32, 64, 128]
        inpl = feature_inplanes[cnn_type.lower()]
        self.conv_lv3 = BasicConv2d(inpl[1], planes[1], 1)
        self.conv_lv4 = BasicConv2d(inpl[2], planes[2], 1)
        self.lv5_aspp1 = BasicConv2d(inpl[3], planes[3], 1)
        for i, rate in enumerate((6, 12, 18)):
            setattr(self, "lv5_aspp%d" % (i + 2), dwBlock(inpl[3], planes[3], 3, dilation=rate))
        self.conv_lv5 = BasicConv2d(planes[3] * 4, planes[3], 1)
        self.conv_last = BasicConv2d(planes[1] + planes[2] + planes[3], last_channel, 3)
        init_weights(self, "kaiming_normal", mode="fan_out")             # before the backbone is attached (:133-135)
        self.features = feature_loader[cnn_type.lower()](name=cnn_type.lower())
        self._planes = planes

    def _emit_from_input(self, plan: Plan, x_src: torch.Tensor, kind: int, n, h, w, taps: Optional[dict] = None):
        (_, _, _), (_, _, _), (c3, h3, w3), (c4, h4, w4), (c5, h5, w5) = self.features._emit_levels(plan, x_src, kind, n, h, w)
        p1, p2, p3 = self._planes[1], self._planes[2], self._planes[3]
        cat5 = plan.alloc(n * h5 * w5, 4 * p3)
        self.lv5_aspp1._emit(plan, c5, n, h5, w5, out=cat5.slot(0, p3), tag="aspp1")
        for i in range(3):
            getattr(self, "lv5_aspp%d" % (i + 2))._emit(plan, c5, n, h5, w5, out=cat5.slot((i + 1) * p3, p3), tag="aspp%d" % (i + 2))
        x5, _, _ = self.conv_lv5._emit(plan, cat5, n, h5, w5, tag="conv_lv5")
        x4, _, _ = self.conv_lv4._emit(plan, c4, n, h4, w4, tag="conv_lv4")
        cat = plan.alloc(n * h3 * w3, p3 + p2 + p1)                       # [x_c5 | x_c4 | x_c3] (model.py:155)
        plan.bilinear(x5, n, h5, w5, p3, cat.slot(0, p3), n, h3, w3, tag="up5")
        plan.bilinear(x4, n, h4, w4, p2, cat.slot(p3, p2), n, h3, w3, tag="up4")
        self.conv_lv3._emit(plan, c3, n, h3, w3, out=cat.slot(p3 + p2, p1), tag="conv_lv3")
        out, _, _ = self.conv_last._emit(plan, cat, n, h3, w3, tag="conv_last")
        if taps is not None:
            taps.update(c3=(c3, h3, w3), c4=(c4, h4, w4), c5=(c5, h5, w5), sfnet=(out, h3, w3))
        return out, h3, w3

    def forward(self, x):
        require_cuda(x, "uavsal_srfnet_aspp")
        n, _, h, w = x.shape

        def build(plan):
            xin = plan.tensor((n, 3, h, w))
            out, ho, wo = self._emit_from_input(plan, xin, 0, n, h, w)
            y = plan.tensor((n, out.c, ho, wo))
            plan.unpack_nchw(out, n, out.c, ho, wo, y)
            plan.named.update(x_in=xin, y_out=y)

        plan = self._cached_plan((x.device, "srf", n, h, w), build)
        plan.named["x_in"].copy_(x)
        plan.launch()
        return plan.named["y_out"].clone()


class spConv(KernelModule):
    """Spatial branch: one dwBlock without residual (model.py:163-171)."""

    def __init__(self, inplanes, planes=256, kernel_size=3, stride=1, expand_ratio=6, dilation=1, res_connect=False):
        super().__init__()
        self.spconv = dwBlock(inplanes, planes, kernel_size, stride, expand_ratio, dilation, res_connect)
        init_weights(self, "kaiming_normal", mode="fan_out")

    def _emit(self, plan, x, n, h, w, tag="sp"):
        return self.spconv._emit(plan, x, n, h, w, tag=tag)

    def forward(self, x):
        return self._forward_nchw(x)


class teConv_sub(KernelModule):
    """Temporal-difference branch (model.py:173-208): 1x1 reduce -> neighbour differences over the call batch
    -> dwBlock -> 1x1 expand."""

    def __init__(self, inplanes, planes=256, time_dims=8, reduction=8, res_connect=False):
        super().__init__()
        self.time_dims = time_dims
        self.res_connect = res_connect and inplanes == planes
        width = planes // reduction
        self.reduce_conv = BasicConv2d(inplanes, width, 1)
        self.sub_conv = dwBlock(2 * width, width, 3, res_connect=False)
        self.last_conv = BasicConv2d(width, planes, 1)
        init_weights(self, "kaiming_normal", mode="fan_out")
        self._width = width

    def _emit(self, plan, x, n, h, w, add: Buf = None, tag="te"):
        """``add`` (optional) is summed onto the result in the last conv's epilogue (STBlock fu_type='sum')."""
        if n < 2:
            raise ValueError("teConv_sub needs at least 2 frames per call (model.py:194 indexes x1[1])")
        x1, _, _ = self.reduce_conv._emit(plan, x, n, h, w, tag=tag + ".reduce")
        d = plan.alloc(n * h * w, 2 * self._width)
        plan.tdiff(x1, n, h * w, self._width, d, tag=tag + ".tdiff")
        s, _, _ = self.sub_conv._emit(plan, d, n, h, w, tag=tag + ".sub")
        res = add
        if self.res_connect:
            if add is not None:
                raise NotImplementedError
            res = x
        out, _, _ = self.last_conv._emit(plan, s, n, h, w, res=res, tag=tag + ".last")
        return out, h, w

    def forward(self, x):
        return self._forward_nchw(x)


class STBlock(KernelModule):
    """x + conv1x1(spConv(x) + teConv_sub(x))  (model.py:210-249, fu_type='sum')."""

    def __init__(self, inplanes, planes=256, time_dims=8, fu_type="sum", res_connect=True, **kwargs):
        super().__init__()
        assert fu_type.lower() in ["sum", "cat"]
        if fu_type.lower() != "sum":
            raise NotImplementedError("fu_type='cat' is not used by UAVSal (model.py:274) and has no kernel plan")
        self.res_connect = res_connect and inplanes == planes
        self.time_dims, self.fu_type, self.inplanes, self.planes = time_dims, fu_type.lower(), inplanes, planes
        self.stconv_sp = spConv(inplanes, planes, res_connect=False)
        self.stconv_te = teConv_sub(inplanes, planes, time_dims, res_connect=False, **kwargs)
        self.stconv_last = BasicConv2d(planes, planes, 1)
        init_weights(self, "kaiming_normal", mode="fan_out")

    def _emit(self, plan, x, n, h, w, out: Buf = None, tag="st"):
        sp, _, _ = self.stconv_sp._emit(plan, x, n, h, w, tag=tag + ".sp")
        summed, _, _ = self.stconv_te._emit(plan, x, n, h, w, add=sp, tag=tag + ".te")      # sp + te (model.py:241)
        y, _, _ = self.stconv_last._emit(plan, summed, n, h, w, out=out, res=x if self.res_connect else None, tag=tag + ".last")
        return y, h, w

    def forward(self, x):
        return self._forward_nchw(x)


class UAVSal(KernelModule):
    """model.py:254-375.  forward(x (N,3,H,W) normalised fp32 | raw uint8, cb=[gauss (N,8,h,w), ob (N,20,h,w)],
    in_state=[h (1,planes,h,w)] | None) -> (out (N,1,h,w) in (0,1), [h_last])."""
    _RNN = "twa"

    def __init__(self, cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1],
                 iosize=[360, 640, 45, 80], planes=256, pre_model_path=""):
        super().__init__()
        self.time_dims = time_dims
        self.sfnet = uavsal_srfnet_aspp(cnn_type, last_channel=planes)
        self.num_stblock = num_stblock
        self.st_layer = nn.Sequential(*[STBlock(planes, planes, time_dims=time_dims, reduction=planes // 32, res_connect=True)
                                        for _ in range(num_stblock)])
        self.fust_layer = nn.Sequential(dwBlock(planes, planes, kernel_size=3))
        self.use_gauss_prior, self.use_ob_prior, self.use_context_prior = bias_type[0], bias_type[1], bias_type[2]
        self.num_cb = int(np.sum(np.array(bias_type) > 0))
        cbp = 64
        if self.use_gauss_prior:
            self.gauss_cb_layer = nn.Sequential(dwBlock(8, cbp, kernel_size=3), dwBlock(cbp, cbp, kernel_size=3))
            init_weights(self.gauss_cb_layer)
        if self.use_ob_prior:
            self.ob_cb_layer = nn.Sequential(dwBlock(20, cbp, kernel_size=3), dwBlock(cbp, cbp, kernel_size=3))
            init_weights(self.ob_cb_layer)
        if self.use_context_prior:
            self.cxt_cb_prior = nn.Sequential(dwBlock(planes, cbp, kernel_size=3, stride=2), dwBlock(cbp, cbp, kernel_size=3, stride=2))
            init_weights(self.cxt_cb_prior)
        if self.num_cb:
            nb = int(np.sum(np.array(bias_type) * cbp))                       # model.py:318
            self.fucb_layer = nn.Sequential(dwBlock(nb, planes // 4, kernel_size=3))
            self.fucbst_layer = nn.Sequential(dwBlock(planes + planes // 4, planes, kernel_size=3))
        _, _, r_out, c_out = iosize
        rnn_cls = ConvTWA if self._RNN == "twa" else ConvLSTM                # UAVSAL_LSTM (model.py:1033) swaps the recurrence
        self.rnn = rnn_cls((r_out, c_out), planes, planes, kernel_size=(3, 3), num_layers=1, batch_first=True, bias=False,
                           return_all_layers=False)
        self.conv_out_st = dwBlock(planes, 1, kernel_size=3)
        for part in (self.st_layer, self.fust_layer, self.conv_out_st):
            init_weights(part, "kaiming_normal", mode="fan_out")
        self._planes, self._iosize, self._bias_type = planes, list(iosize), list(bias_type)
        if os.path.exists(pre_model_path):
            print("Load pre-trained weights")
            self.load_reference(pre_model_path, strict=False)                  # model.py:339

    def load_reference(self, path, strict: bool = True):
        """``self.load_state_dict(torch.load(path).state_dict())`` (Demo_Test.py:39) for the reference's whole-module pickles,
        without its classes (or the torchvision version it was saved with) being importable: see checkpoint.py."""
        from . import checkpoint
        return checkpoint.load_into(self, path, strict=strict)

    # -----------------------------------------------------------------------------------------------
    def build_plan(self, plan: Plan, n: int, h: int, w: int, x_kind: int = 0, post_hw=None, taps: bool = False,
                   cb_shared: bool = False, stage: str = "all", group: int = 0, clips: int = 1):
        """Emit the whole forward for a call of n frames of (h, w) pixels.
        x_kind: 0 fp32 NCHW normalised, 1 uint8 NCHW raw, 2 uint8 NHWC raw.  post_hw=(H,W) adds the uint8 post-process.
        cb_shared: cb tensors hold ONE frame that is broadcast to all n (Demo_Test's np.repeat'ed priors).
        stage: "all" (a reference call), "sfnet" (only the per-frame SRF-Net, any n: frames are independent there, so a
        runner may batch a whole clip), "head" (everything after the SRF-Net, fed from the arena input ``sf_in``).
        group: > 0 = the n frames are CONSECUTIVE reference calls of `group` frames each (the last may be shorter) emitted as
        one plan: the only call-granular operations - the temporal differences' mirrored edges (model.py:194-198) and the
        context prior's repeat interleave (model.py:361) - are told the call size; the ConvTWA state simply runs through.
        clips: the n frames are `clips` independent clips of n/clips frames each (clip-major); everything up to the ConvTWA is
        per frame / per call group anyway, the ConvTWA runs the clips as a batch of sequences (own state each)."""
        planes, T = self._planes, self.time_dims
        assert stage in ("all", "sfnet", "head")
        if clips > 1 and (n % clips or not group or (n // clips) % group):
            raise ValueError("batched clips need n divisible by clips and clips made of whole calls (n=%d clips=%d group=%d)" % (n, clips, group))
        if group:
            if group % T or (n % group) % T:
                raise ValueError("call size %d / batch %d must be multiples of time_dims=%d" % (group, n, T))
            plan.call_group = group
        if n % T and stage != "sfnet":
            raise ValueError("call batch %d is not a multiple of time_dims=%d (model.py:356-357)" % (n, T))
        tp = {} if taps else None
        if stage == "head":
            mh, mw = self._iosize[2], self._iosize[3]
            if (out_size(out_size(out_size(h, 2), 2), 2), out_size(out_size(out_size(w, 2), 2), 2)) != (mh, mw):
                mh, mw = out_size(out_size(out_size(h, 2), 2), 2), out_size(out_size(out_size(w, 2), 2), 2)
            x = plan.alloc(n * mh * mw, planes)
            x_in = None
            plan.named["sf_in"] = x
        else:
            x_in = plan.tensor((n, 3, h, w) if x_kind < 2 else (n, h, w, 3), torch.float32 if x_kind == 0 else torch.uint8)
            x, mh, mw = self.sfnet._emit_from_input(plan, x_in, x_kind, n, h, w, taps=tp)
            if stage == "sfnet":
                plan.named.update(x_in=x_in, sf_out=x, map_hw=(mh, mw))
                return plan
        for i, blk in enumerate(self.st_layer):
            x, _, _ = blk._emit(plan, x, n, mh, mw, tag="st%d" % i)
            if taps:
                tp["st_layer.%d" % i] = (x, mh, mw)
        rows = n * mh * mw
        named = dict(x_in=x_in)
        if self.num_cb:
            q = planes // 4
            cat2 = plan.alloc(rows, planes + q)                              # [x | x_cb] (model.py:365)
            x, _, _ = self.fust_layer[0]._emit(plan, x, n, mh, mw, out=cat2.slot(0, planes), tag="fust")
            nb = 64 * self.num_cb
            cat1 = plan.alloc(rows, nb)                                      # [gauss | ob | cxt] (model.py:363)
            off = 0
            ncb = 1 if cb_shared else n
            if self.use_gauss_prior:
                g_in = plan.tensor((ncb, 8, mh, mw))
                gb = plan.alloc(ncb * mh * mw, 8)
                plan.pack_nchw(g_in, ncb, 8, mh, mw, gb, tag="cb_gauss.pack")
                y, _, _ = self.gauss_cb_layer[0]._emit(plan, gb, ncb, mh, mw, tag="gauss0")
                if cb_shared:
                    y, _, _ = self.gauss_cb_layer[1]._emit(plan, y, ncb, mh, mw, tag="gauss1")
                    plan.bilinear(y, 1, mh, mw, 64, cat1.slot(off, 64), n, mh, mw, tag="gauss.bcast")
                else:
                    self.gauss_cb_layer[1]._emit(plan, y, ncb, mh, mw, out=cat1.slot(off, 64), tag="gauss1")
                named["cb_gauss_in"] = g_in
                if taps:
                    tp["cb_gauss"] = (cat1.slot(off, 64), mh, mw)
                off += 64
            if self.use_ob_prior:
                o_in = plan.tensor((ncb, 20, mh, mw))
                ob = plan.alloc(ncb * mh * mw, 24)
                ob.c = 24
                plan.pack_nchw(o_in, ncb, 20, mh, mw, ob, tag="cb_ob.pack")
                y, _, _ = self.ob_cb_layer[0]._emit(plan, ob, ncb, mh, mw, tag="ob0")
                if cb_shared:
                    y, _, _ = self.ob_cb_layer[1]._emit(plan, y, ncb, mh, mw, tag="ob1")
                    plan.bilinear(y, 1, mh, mw, 64, cat1.slot(off, 64), n, mh, mw, tag="ob.bcast")
                else:
                    self.ob_cb_layer[1]._emit(plan, y, ncb, mh, mw, out=cat1.slot(off, 64), tag="ob1")
                named["cb_ob_in"] = o_in
                if taps:
                    tp["cb_ob"] = (cat1.slot(off, 64), mh, mw)
                off += 64
            if self.use_context_prior:
                b = n // T
                s = plan.alloc(b * mh * mw, planes)
                plan.ctx_sum(x, b, T, mh * mw, planes, s, tag="cxt.sum")
                y, h1, w1 = self.cxt_cb_prior[0]._emit(plan, s, b, mh, mw, tag="cxt0")
                y, h2, w2 = self.cxt_cb_prior[1]._emit(plan, y, b, h1, w1, tag="cxt1")
                # upsample (align_corners) + repeat(T): frame i reads chunk i % b (model.py:360-361, quirk Q3)
                plan.bilinear(y, b, h2, w2, 64, cat1.slot(off, 64), n, mh, mw, tag="cxt.up",
                              src_group=(group // T) if group else 0, dst_group=group)
                off += 64
            self.fucb_layer[0]._emit(plan, cat1, n, mh, mw, out=cat2.slot(planes, q), tag="fucb")
            x, _, _ = self.fucbst_layer[0]._emit(plan, cat2, n, mh, mw, tag="fucbst")
            if taps:
                tp.update(fust=(cat2.slot(0, planes), mh, mw), fucb=(cat2.slot(planes, q), mh, mw), fucbst=(x, mh, mw))
        else:
            x, _, _ = self.fust_layer[0]._emit(plan, x, n, mh, mw, tag="fust")
        # temporal weighted average over the call's frames, batch 1 (model.py:367-370).  Everything from here on depends on
        # the previous call's hidden state: it is the plan's "back" part (runner.ClipRunner overlaps it with the next front)
        plan.mark_split()
        h_in = plan.tensor((clips, planes, mh, mw))
        hb = plan.alloc(clips * mh * mw, planes)
        plan.pack_nchw(h_in, clips, planes, mh, mw, hb, tag="state.pack")
        seq = plan.alloc(rows, planes)
        cell = self.rnn.cell_list[0]
        if isinstance(self.rnn, ConvLSTM):
            # UAVSAL_LSTM (model.py:1065-1068): 4-gate ConvLSTM over the call's frames; the cell state lives in the plan as NHWC fp32
            c_state = plan.tensor((clips, mh * mw, planes))
            plan.lstm(x, hb, c_state, clips, n // clips, mh, mw, planes, planes, cell.wspec(), None, seq, tag="rnn")
            named.update(c_state=c_state)
        else:
            emit_twa(plan, cell, x, hb, seq, clips, n // clips, mh, mw)
        h_out = plan.tensor((clips, planes, mh, mw))
        for ci in range(clips):
            last = seq.at_row(((ci + 1) * (n // clips) - 1) * mh * mw)
            plan.unpack_nchw(last, 1, planes, mh, mw, h_out[ci:ci + 1], tag="state.unpack")
        # readout: expand + dw (dwBlock 256->1), then the 1536->1 project + BN + sigmoid as a dot product (model.py:372-373)
        ro = self.conv_out_st
        e, _, _ = ro.conv[0]._emit(plan, seq, n, mh, mw, tag="readout.expand", out_fmt=plan.hidden_fmt(ro.geom[2]))
        out = plan.tensor((n, 1, mh, mw))
        if e.plain and getattr(plan, "fuse_readout", True):
            plan.dw_dot_sigmoid(e, n, mh, mw, e.c, ro.conv[1].wspec(), None, ro.project_wspec(), None, out, tag="readout.dw+dot")
        else:
            d, _, _ = ro.conv[1]._emit(plan, e, n, mh, mw, tag="readout.dw")
            plan.dot_sigmoid(d, rows, e.c, ro.project_wspec(), None, out, tag="readout.dot")
        named.update(h_in=h_in, h_out=h_out, out=out, map_hw=(mh, mw))
        if taps:
            tp["rnn"] = (seq, mh, mw)
            named["taps"] = tp
        if post_hw is not None:
            u8 = plan.tensor((n, post_hw[0], post_hw[1]), torch.uint8)
            plan.post_u8(out, n, mh, mw, post_hw[0], post_hw[1], u8, tag="post")
            named["out_u8"] = u8
        plan.named.update(named)
        return plan

    def get_plan(self, dev, n, h, w, x_kind=0, post_hw=None, taps=False, cb_shared=False, slot=0, stage="all", group=0, clips=1) -> Plan:
        """``slot`` selects one of several independent plan instances (own arena) so that calls can be in flight together."""
        key = (torch.device(dev), "uavsal", n, h, w, x_kind, post_hw, taps, cb_shared, slot, stage, group, clips)
        return self._cached_plan(key, lambda plan: self.build_plan(plan, n, h, w, x_kind, post_hw, taps, cb_shared, stage, group, clips))

    def forward(self, x, cb, in_state):
        require_cuda(x, "UAVSal")
        n, _, h, w = x.shape
        kind = 1 if x.dtype == torch.uint8 else 0
        plan = self.get_plan(x.device, n, h, w, kind)
        nm = plan.named
        nm["x_in"].copy_(x)
        mh, mw = nm["map_hw"]
        if self.use_gauss_prior:
            nm["cb_gauss_in"].copy_(cb[0])
        if self.use_ob_prior:
            nm["cb_ob_in"].copy_(cb[1])
        if in_state is None:
            nm["h_in"].zero_()                                   # reference: init_hidden zeros (model_convlstm.py:294)
        else:
            nm["h_in"].copy_(in_state[0])
        plan.launch()
        return nm["out"].clone(), [nm["h_out"].clone()]


class UAVSAL_LSTM(UAVSal):
    """model.py:960-1076 (the Table-V ablation): UAVSal with the 4-gate ``ConvLSTM`` as the recurrence.
    forward(x, cb, in_state=[[h, c]] | None) -> (out (N,1,h,w), [h_last, c_last]) as ``ConvLSTM.forward`` returns them
    (model_convlstm.py:196-218: one [h, c] pair per layer in, the last layer's pair out)."""
    _RNN = "lstm"

    def __init__(self, cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1],
                 iosize=[360, 640, 45, 80], planes=256, pre_model_path=""):
        super().__init__(cnn_type, time_dims, num_stblock, bias_type, iosize, planes, pre_model_path)

    def forward(self, x, cb, in_state):
        require_cuda(x, "UAVSAL_LSTM")
        n, _, h, w = x.shape
        plan = self.get_plan(x.device, n, h, w, 1 if x.dtype == torch.uint8 else 0)
        nm = plan.named
        nm["x_in"].copy_(x)
        mh, mw = nm["map_hw"]
        if self.use_gauss_prior:
            nm["cb_gauss_in"].copy_(cb[0])
        if self.use_ob_prior:
            nm["cb_ob_in"].copy_(cb[1])
        if in_state is None:
            nm["h_in"].zero_()
            nm["c_state"].zero_()
        else:
            h0, c0 = in_state[0]
            nm["h_in"].copy_(h0)
            nm["c_state"].copy_(c0.permute(0, 2, 3, 1).reshape(nm["c_state"].shape))      # boundary layout change (NCHW -> NHWC)
        plan.launch()
        c_last = nm["c_state"].view(1, mh, mw, self._planes).permute(0, 3, 1, 2).contiguous()
        return nm["out"].clone(), [nm["h_out"].clone(), c_last]

