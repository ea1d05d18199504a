"""Plan builder + executor: turns the module tree into a flat list of C-ABI kernel calls over a
pre-allocated activation arena (NHWC, split-bf16 planes — see include/uavsal_b200.h).

Nothing here computes on the CPU: a plan can be *built* on any device (the CPU test-suite checks shapes,
packing and the op list that way) but ``Plan.run`` requires CUDA tensors and the sm_100a library.
"""
from __future__ import annotations

import ctypes
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _ext

F_RELU6, F_RESIDUAL, F_SIGMOID, F_OUT_F32, F_RELU, F_OUT_Q16, F_HID_Q16 = 1, 2, 4, 8, 16, 32, 64
TERMS_GEN1 = 0x100        # UAVSAL_TERMS_GEN1
PLANE_F32 = -1            # UAVSAL_PLANE_F32: the activation is plain fp32 rows, not split-bf16 planes
PLANE_Q16 = -2            # UAVSAL_PLANE_Q16: uint16 fixed-point rows of a ReLU6 output, q = rne(v * 65535 / 6)
FMT_SPLIT, FMT_F32, FMT_Q16 = 0, 1, 2
Q16_HIDDEN_MIN = 1152     # hidden tensors at least this wide travel as q16 rows (see Plan.hidden_fmt)
BN_EPS = 1e-5


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def out_size(n: int, stride: int) -> int:
    """3x3, padding = dilation, stride 1|2 (model.py:67-69)."""
    return n if stride == 1 else (n - 1) // 2 + 1


class _Alloc:
    """One arena allocation: size, [first, last] op index that touches it, offset once the layout is solved."""
    __slots__ = ("idx", "nbytes", "first", "last", "pinned", "off")

    def __init__(self, idx, nbytes):
        self.idx, self.nbytes, self.first, self.last, self.pinned, self.off = idx, nbytes, None, None, False, None


class Buf:
    """A (rows x c) activation living in a (2, rows, ld) bf16 tensor at channel offset ``off`` (``fmt`` FMT_SPLIT) — or in
    plain rows: a (rows, ld) fp32 tensor (FMT_F32) or uint16 fixed point (FMT_Q16), the two forms of the hidden tensor between
    an expand conv and its depthwise conv.  ``root`` ties every slot / row view to the arena allocation it lives in
    (liveness tracking of two-pass plans)."""

    __slots__ = ("t", "rows", "c", "ld", "off", "fmt", "root", "plan", "base")

    def __init__(self, t, rows: int, c: int, ld: int, off: int = 0, fmt: int = FMT_SPLIT, root=None, plan=None, base: int = 0):
        self.t, self.rows, self.c, self.ld, self.off, self.fmt = t, rows, c, ld, off, int(fmt)
        self.root, self.plan, self.base = root, plan, base

    @property
    def f32(self) -> bool:
        return self.fmt == FMT_F32

    @property
    def q16(self) -> bool:
        return self.fmt == FMT_Q16

    @property
    def plain(self) -> bool:
        """Plain rows (fp32 | q16): written by the expand GEMM, read by the depthwise kernels only."""
        return self.fmt != FMT_SPLIT

    def _touch(self):
        if self.root is not None:
            self.plan._touched.add(self.root)

    @property
    def ptr(self) -> int:
        self._touch()
        base = self.t.data_ptr() if self.t is not None else self.base
        return base + (4 if self.f32 else 2) * self.off

    @property
    def plane(self) -> int:
        return PLANE_F32 if self.f32 else PLANE_Q16 if self.q16 else self.rows * self.ld

    def act(self) -> Tuple[int, int, int]:
        return (self.ptr, self.plane, self.ld)

    def slot(self, off: int, c: int) -> "Buf":
        assert off % 8 == 0 and off + c <= self.ld
        return Buf(self.t, self.rows, c, self.ld, self.off + off, self.fmt, self.root, self.plan, self.base)

    def at_row(self, row: int) -> "Buf":
        """The same buffer seen from row ``row`` on (same plane distance and pitch: used for the last frame of a sequence)."""
        return Buf(self.t, self.rows, self.c, self.ld, self.off + row * self.ld, self.fmt, self.root, self.plan, self.base)

    def to_float(self) -> torch.Tensor:
        """fp32 (rows, c) reconstruction hi + lo (debug / tests)."""
        if self.q16:
            v = self.t.view(torch.int16).to(torch.int32).bitwise_and(0xFFFF).float() * (6.0 / 65535.0)
        else:
            v = self.t if self.f32 else self.t[0].float() + self.t[1].float()
        return v[:, self.off:self.off + self.c]


NULL_ACT = (0, 0, 0)


# ---------------------------------------------------------------------------------------------------
# weight preparation (BN folding, hi/lo split, layouts)
# ---------------------------------------------------------------------------------------------------
def fold_bn(w: torch.Tensor, bn) -> Tuple[torch.Tensor, torch.Tensor]:
    """conv(no bias) + BatchNorm2d(eval) -> (w', b') with w' = w*g/sqrt(var+eps), b' = beta - mean*g/sqrt(var+eps)
    (model.py:69-70, 94-95; eps = 1e-5)."""
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    bias = bn.bias.detach().float() - bn.running_mean.detach().float() * scale
    return w.detach().float() * scale.view(-1, 1, 1, 1), bias


def split_bf16(w: torch.Tensor) -> torch.Tensor:
    """fp32 (...)-> bf16 (2, ...) planes: hi = bf16(w), lo = bf16(w - hi)."""
    hi = w.to(torch.bfloat16)
    lo = (w - hi.float()).to(torch.bfloat16)
    return torch.stack([hi, lo], 0).contiguous()


def pack_pw_tc(w2d: torch.Tensor, kpad: int) -> torch.Tensor:
    n, k = w2d.shape
    full = torch.zeros((n, kpad), dtype=torch.float32, device=w2d.device)
    full[:, :k] = w2d
    return split_bf16(full)


def pack_pw_simt(w2d: torch.Tensor, kpad: int) -> torch.Tensor:
    n, k = w2d.shape
    full = torch.zeros((kpad, n), dtype=torch.float32, device=w2d.device)
    full[:k] = w2d.t()
    return full.contiguous()


def pack_dw(w: torch.Tensor) -> torch.Tensor:
    c = w.shape[0]
    return w.reshape(c, 9).t().contiguous()            # [9][C]


def conv3x3_as_2d(w: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, 3, 3) -> (Cout, 9*Cin) with k = (ky*3+kx)*Cin + ci."""
    cout, cin = w.shape[:2]
    return w.permute(0, 2, 3, 1).reshape(cout, 9 * cin).contiguous()


def interleave_gates(w: torch.Tensor, ch: int) -> torch.Tensor:
    """ConvLSTM rows g*ch + c (i,f,o,g blocks, model_convlstm.py:117) -> 4*c + g."""
    return w.reshape(4, ch, *w.shape[1:]).transpose(0, 1).reshape(4 * ch, *w.shape[1:]).contiguous()


W_ROWS_SPLIT, W_ROWS_F32, W_COLS_F32 = 0, 1, 2        # UAVSAL_W_* layouts of uavsal_pack_weights


class W:
    """A conv layer's parameters as the module holds them: weight (cout, cin_per_group, kh, kw) [or 2-D], the BatchNorm2d that
    follows it (eval statistics are folded in, model.py:69-70, 94-95) and / or the conv's own bias.  ``owner`` (the nn.Conv2d)
    carries the cache of packed forms, keyed on the layout and validated against every tensor's (data_ptr, _version)."""

    def __init__(self, weight: torch.Tensor, bn=None, bias: Optional[torch.Tensor] = None, owner=None):
        self.weight, self.bn, self.bias, self.owner = weight, bn, bias, owner
        self.cout = weight.shape[0]
        self.cin = weight.shape[1]
        self.taps = int(weight.numel() // (self.cout * self.cin))

    def tensors(self):
        ts = [self.weight]
        if self.bn is not None:
            ts += [self.bn.weight, self.bn.bias, self.bn.running_mean, self.bn.running_var]
        if self.bias is not None:
            ts.append(self.bias)
        return ts

    def signature(self):
        return tuple((id(t), t.data_ptr(), t._version) for t in self.tensors()) + ((self.bn.eps,) if self.bn is not None else ())

    # torch restatement of uavsal_pack_weights: used when a plan is built without a device (structure tests on the CPU) and as
    # the oracle of the kernel's parity test
    def pack_reference(self, layout: int, n_pad: int, k_pad: int, gates: int = 1):
        w = self.weight.detach().float().reshape(self.cout, self.cin, self.taps)
        b = self.bias.detach().float() if self.bias is not None else torch.zeros(self.cout, device=w.device)
        if self.bn is not None:
            scale = self.bn.weight.detach().float() / torch.sqrt(self.bn.running_var.detach().float() + self.bn.eps)
            w = w * scale.view(-1, 1, 1)
            b = self.bn.bias.detach().float() + (b - self.bn.running_mean.detach().float()) * scale
        if gates > 1:
            w, b = interleave_gates(w, self.cout // gates), interleave_gates(b, self.cout // gates)
        full = torch.zeros((n_pad, k_pad), dtype=torch.float32, device=w.device)
        full[:self.cout, :self.taps * self.cin] = w.permute(0, 2, 1).reshape(self.cout, self.taps * self.cin)      # k = tap * cin + ci
        bias = torch.zeros((n_pad,), dtype=torch.float32, device=w.device)
        bias[:self.cout] = b
        if layout == W_ROWS_SPLIT:
            return split_bf16(full), bias
        if layout == W_COLS_F32:
            return full.t().contiguous(), bias
        return full, bias


_DUMMY = None


def _dummy():
    global _DUMMY
    if _DUMMY is None:
        _DUMMY = torch.zeros(16)
    return _DUMMY


def invalidate_packed(module: torch.nn.Module):
    """Drop every cached packed weight below ``module`` (needed after editing parameters in a way autograd's version counter
    does not see, e.g. ``p.data.copy_(...)`` or a write through a raw pointer)."""
    for m in module.modules():
        m.__dict__.pop("_uavsal_packed", None)


# ---------------------------------------------------------------------------------------------------
# plan
# ---------------------------------------------------------------------------------------------------
@dataclass
class Op:
    name: str
    fn: Callable
    args: tuple
    tag: str = ""


class Plan:
    """mode "direct" (default): every allocation is its own tensor - what tests and tools use when they drive single kernels.
    Modules build their cached plans in two passes (``Plan.build``): a "measure" pass records every allocation and the ops
    that touch it, then the real pass places all of them in ONE arena allocation, buffers with disjoint lifetimes sharing
    memory (the 6x hidden tensors of the inverted-residual blocks dominate: 36 GB -> a few GB for a 120-frame plan)."""

    ALIGN = 1024

    def __init__(self, device: torch.device, terms: int = 3, engine: str = "tc", mode: str = "direct", layout=None):
        assert terms in (1, 3) and engine in ("tc", "tc1", "simt") and mode in ("direct", "measure", "arena")
        self.device = torch.device(device)
        self.terms = terms
        self.engine = engine
        self.terms_arg = terms | (TERMS_GEN1 if engine == "tc1" else 0)      # the kernel generation travels with every call
        self.mode = mode
        self.ops: List[Op] = []
        self.keep: List[torch.Tensor] = []       # packed weights / scratch kept alive with the plan
        self.named: Dict[str, object] = {}       # name -> Buf / tensor (inputs, outputs, debug taps)
        self.graph = None
        self.arena_bytes = 0
        self._allocs: List[_Alloc] = []
        self._touched = set()
        self._layout = layout
        self._arena = None
        if mode == "arena":
            offsets, total = layout
            self._arena = torch.zeros((max(total, 16),), dtype=torch.uint8, device=self.device)
            self.arena_bytes = total
            self.keep.append(self._arena)

    # ---- two-pass construction ----
    @classmethod
    def build(cls, device, terms: int, engine: str, builder: Callable[["Plan"], None]) -> "Plan":
        device = torch.device(device)
        if device.type != "cuda":
            plan = cls(device, terms, engine)
            builder(plan)
            return plan
        probe = cls(device, terms, engine, mode="measure")
        builder(probe)
        plan = cls(device, terms, engine, mode="arena", layout=probe.solve_layout())
        builder(plan)
        assert len(plan._allocs) == len(probe._allocs), "plan builders must be deterministic"
        return plan

    def solve_layout(self):
        """Offsets for the recorded allocations: persistent ones (module-boundary tensors, anything reachable from ``named``,
        never-touched ones) first, the rest first-fit in allocation order with memory returned after the last op that
        touches it.  Two buffers share memory only if one's last op strictly precedes the other's first."""
        def reach(o, out):
            if isinstance(o, Buf):
                if o.root is not None:
                    out.add(o.root)
            elif isinstance(o, dict):
                for v in o.values():
                    reach(v, out)
            elif isinstance(o, (list, tuple)):
                for v in o:
                    reach(v, out)
        pinned = set()
        reach(self.named, pinned)
        for a in self._allocs:
            if a in pinned or a.first is None:
                a.pinned = True
        up = lambda n: (n + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        off = 0
        for a in self._allocs:
            if a.pinned:
                a.off = off
                off += up(a.nbytes)
        base = off
        live = []                                   # (off, end, last) of transient buffers currently holding memory
        peak = base
        for a in self._allocs:
            if a.pinned:
                continue
            live = [x for x in live if x[2] >= a.first]
            live.sort()
            pos = base
            for lo, hi, _ in live:
                if lo - pos >= a.nbytes:
                    break
                pos = max(pos, hi)
            a.off = pos
            live.append((pos, pos + up(a.nbytes), a.last))
            peak = max(peak, pos + up(a.nbytes))
        return [a.off for a in self._allocs], peak

    # ---- allocation ----
    def _new(self, nbytes: int, persistent: bool = False):
        a = _Alloc(len(self._allocs), nbytes)
        a.pinned = persistent
        self._allocs.append(a)
        return a

    def _carve(self, a: _Alloc, dtype, shape):
        off = self._layout[0][a.idx]
        return self._arena[off:off + a.nbytes].view(dtype).view(shape)

    def alloc(self, rows: int, c: int, ld: Optional[int] = None) -> Buf:
        ld = _pad8(c) if ld is None else ld
        if self.mode == "direct":
            t = torch.zeros((2, rows, ld), dtype=torch.bfloat16, device=self.device)
            self.arena_bytes += t.numel() * 2
            self.keep.append(t)
            return Buf(t, rows, c, ld)
        a = self._new(2 * rows * ld * 2)
        t = self._carve(a, torch.bfloat16, (2, rows, ld)) if self.mode == "arena" else None
        return Buf(t, rows, c, ld, 0, False, a, self, 4096 * (a.idx + 1))

    def alloc_f32(self, rows: int, c: int) -> Buf:
        """fp32 rows (only the persistent tcgen05 GEMM / the stem write them, only the TMA depthwise kernel reads them)."""
        ld = _pad8(c)
        if self.mode == "direct":
            t = torch.zeros((rows, ld), dtype=torch.float32, device=self.device)
            self.arena_bytes += t.numel() * 4
            self.keep.append(t)
            return Buf(t, rows, c, ld, 0, FMT_F32)
        a = self._new(rows * ld * 4)
        t = self._carve(a, torch.float32, (rows, ld)) if self.mode == "arena" else None
        return Buf(t, rows, c, ld, 0, FMT_F32, a, self, 4096 * (a.idx + 1))

    def alloc_q16(self, rows: int, c: int) -> Buf:
        """uint16 fixed-point rows (q = rne(v * 65535 / 6) of a ReLU6 output; only the persistent tcgen05 GEMM writes them, only
        the depthwise kernels read them).  Held as an int16 tensor (same bits)."""
        ld = _pad8(c)
        if self.mode == "direct":
            t = torch.zeros((rows, ld), dtype=torch.int16, device=self.device)
            self.arena_bytes += t.numel() * 2
            self.keep.append(t)
            return Buf(t, rows, c, ld, 0, FMT_Q16)
        a = self._new(rows * ld * 2)
        t = self._carve(a, torch.int16, (rows, ld)) if self.mode == "arena" else None
        return Buf(t, rows, c, ld, 0, FMT_Q16, a, self, 4096 * (a.idx + 1))

    def alloc_hidden(self, rows: int, c: int, fmt: int) -> Buf:
        return self.alloc_q16(rows, c) if fmt == FMT_Q16 else self.alloc_f32(rows, c) if fmt == FMT_F32 else self.alloc(rows, c)

    @property
    def f32_hidden(self) -> bool:
        return self.engine == "tc"

    def hidden_fmt(self, hidden: int, dilation: int = 1, hw: int = 0) -> int:
        """Storage of the tensor between a dwBlock's expand conv and its depthwise conv (model.py:90-92).  The tcgen05 engine
        keeps it in plain rows for the TMA depthwise kernels (dilation 1): fp32, or - for the widest blocks (hidden >=
        Q16_HIDDEN_MIN: the 256 -> 1536 class, whose 2.65 GB hidden tensors dominate the plan's HBM traffic) - 16-bit fixed point
        of the ReLU6 output (|error| <= 4.6e-5; measured on config #2: +3.6e-5 max-abs on the saliency map).  ``hidden_q16 = False``
        on the plan (or UAVSAL_HIDDEN_Q16=0 in the environment, for A/B runs) keeps everything in fp32."""
        if not self.f32_hidden:
            return FMT_SPLIT
        if not hasattr(self, "hidden_q16"):
            self.hidden_q16 = os.environ.get("UAVSAL_HIDDEN_Q16", "1") != "0"
        if dilation != 1:
            # dilated depthwise convs (the ASPP branches) read plain rows only through the whole-image kernel for small maps
            # (dw_tma.cu dw3x3_img_kernel: two q16 images of 64 channels in shared memory)
            return FMT_Q16 if (hidden >= Q16_HIDDEN_MIN and getattr(self, "hidden_q16", True) and 0 < hw <= 768) else FMT_SPLIT
        if hidden >= Q16_HIDDEN_MIN and getattr(self, "hidden_q16", True):
            return FMT_Q16
        return FMT_F32

    def tensor(self, shape, dtype=torch.float32) -> torch.Tensor:
        """A module-boundary tensor / workspace: persistent (never shares memory), zero-initialised."""
        if self.mode == "direct":
            t = torch.zeros(shape, dtype=dtype, device=self.device)
            self.arena_bytes += t.numel() * t.element_size()
            self.keep.append(t)
            return t
        if self.mode == "measure":
            t = torch.empty(shape, dtype=dtype)                  # host placeholder (pages are never touched): shapes / slicing only
            self._new(t.numel() * t.element_size(), persistent=True)
            return t
        n = 1
        for d in shape:
            n *= int(d)
        a = self._new(n * torch.empty((), dtype=dtype).element_size(), persistent=True)
        return self._carve(a, dtype, tuple(int(d) for d in shape))

    def hold(self, t: torch.Tensor) -> torch.Tensor:
        if self.mode == "measure":
            return t
        t = t.to(self.device).contiguous()
        self.keep.append(t)
        return t

    def packed(self, ws: W, layout: int, n_pad: Optional[int] = None, k_pad: Optional[int] = None, gates: int = 1, host: bool = False):
        """(weight, bias) of a conv layer in the form a kernel reads (BN folded, layout changed, bf16 hi / lo split): one
        uavsal_pack_weights launch, cached on the owning module.  host=True: copies in host memory (kernel-parameter weights)."""
        n_pad = ws.cout if n_pad is None else n_pad
        k_pad = ws.cin * ws.taps if k_pad is None else k_pad
        if self.mode == "measure":
            return _dummy(), _dummy()
        key = (layout, n_pad, k_pad, gates, host, str(self.device))
        cache = ws.owner.__dict__.setdefault("_uavsal_packed", {}) if ws.owner is not None else None
        sig = ws.signature()
        hit = cache.get(key) if cache is not None else None
        if hit is not None and hit[0] == sig:
            wt, b = hit[1], hit[2]
        else:
            if self.device.type == "cuda":
                ts = [t.detach().to(self.device, torch.float32).contiguous() for t in ws.tensors()]
                wsrc = ts[0]
                bn = ts[1:5] if ws.bn is not None else [None] * 4
                cb = ts[-1] if ws.bias is not None else None
                wt = torch.empty((2, n_pad, k_pad), dtype=torch.bfloat16, device=self.device) if layout == W_ROWS_SPLIT else \
                    torch.empty((k_pad, n_pad) if layout == W_COLS_F32 else (n_pad, k_pad), dtype=torch.float32, device=self.device)
                b = torch.empty((n_pad,), dtype=torch.float32, device=self.device)
                ptr = lambda t: t.data_ptr() if t is not None else None
                _ext.call("uavsal_pack_weights", wsrc.data_ptr(), ws.cout, ws.cin, ws.taps, ptr(bn[0]), ptr(bn[1]), ptr(bn[2]), ptr(bn[3]),
                          float(ws.bn.eps) if ws.bn is not None else 0.0, ptr(cb), gates, layout, n_pad, k_pad, wt.data_ptr(), b.data_ptr(),
                          ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
                for t in ts:
                    t.record_stream(torch.cuda.current_stream(self.device))
            else:
                wt, b = ws.pack_reference(layout, n_pad, k_pad, gates)
            if host:
                wt, b = wt.cpu().contiguous(), b.cpu().contiguous()
            if cache is not None:
                cache[key] = (sig, wt, b)
        self.keep += [wt, b]
        return wt, b

    def _add(self, name: str, args: Sequence, tag: str = ""):
        fn = getattr(_ext.load(), name) if (self.device.type == "cuda" and self.mode != "measure") else None
        idx = len(self.ops)
        for a in self._touched:
            if a.first is None:
                a.first = idx
            a.last = idx
        self._touched.clear()
        self.ops.append(Op(name, fn, tuple(args), tag))

    # ---- ops ----
    def pack_nchw(self, src: torch.Tensor, n, c, h, w, dst: Buf, tag=""):
        self._add("uavsal_pack_nchw_f32", (src.data_ptr(), n, c, h, w, *dst.act(), _pad8(c)), tag)

    def unpack_nchw(self, src: Buf, n, c, h, w, dst: torch.Tensor, tag=""):
        self._add("uavsal_unpack_nchw_f32", (*src.act(), n, c, h, w, dst.data_ptr()), tag)

    def stem(self, x: torch.Tensor, kind: int, n, h, w, wgt, bias, out: Buf, tag=""):
        """wgt: W (3->32, 3x3, BN) or folded fp32 [3][3][3][32] (ky, kx, cin, cout) + bias [32].  The tcgen05 engines hand the 3.5 KB
        of weights over as HOST arrays: they travel in the kernel parameters and reach the FMAs as warp-uniform constant-bank operands."""
        host = self.engine != "simt" and ((isinstance(wgt, W) and (wgt.cout, wgt.cin, wgt.taps) == (32, 3, 9)) or
                                          (not isinstance(wgt, W) and tuple(wgt.shape) == (3, 3, 3, 32)))
        if isinstance(wgt, W):
            wd, bd = self.packed(wgt, W_COLS_F32, host=host)
        elif host:
            wd, bd = wgt.detach().float().cpu().contiguous(), bias.detach().float().cpu().contiguous()
            self.keep += [wd, bd]
        else:
            wd, bd = self.hold(wgt), self.hold(bias)
        self._add("uavsal_stem_conv3x3s2_hw" if host else "uavsal_stem_conv3x3s2",
                  (x.data_ptr(), kind, n, h, w, wd.data_ptr(), bd.data_ptr(), *out.act()), tag)

    def _dw_weights(self, wgt, bias):
        """W (depthwise conv + BN) or pre-folded ([9][C] fp32, bias) -> device tensors."""
        if isinstance(wgt, W):
            return self.packed(wgt, W_COLS_F32)
        return self.hold(wgt), self.hold(bias.float())

    def dw(self, x: Buf, n, h, w, c, stride, dil, wgt, bias, relu6, out: Buf, tag=""):
        wd, bd = self._dw_weights(wgt, bias)
        self._add("uavsal_dw3x3", (*x.act(), n, h, w, c, stride, dil, wd.data_ptr(), bd.data_ptr(), int(relu6), *out.act()), tag)

    def expdw(self, x: Buf, n, h, w, w1, b1, stride: int, wd, bd, out: Buf, tag=""):
        """Fused 1x1 expand + BN + ReLU6 -> depthwise 3x3 + BN + ReLU6 (cin <= 32).  w1: W or folded fp32 (hidden, cin_logical) + b1;
        wd: W or folded [9][hidden] + bd."""
        cin = x.c
        if isinstance(w1, W):
            hidden, k = w1.cout, w1.cin
        else:
            hidden, k = w1.shape
        assert not x.plain and cin % 8 == 0 and k <= cin <= 32 and hidden % 8 == 0
        kp = (cin + 15) // 16 * 16
        hp = (hidden + 63) // 64 * 64
        if isinstance(w1, W):
            wp, bp = self.packed(w1, W_ROWS_SPLIT, hp, kp)
        else:
            full = torch.zeros((hp, kp), dtype=torch.float32, device=w1.device)
            full[:hidden, :k] = w1
            bfull = torch.zeros((hp,), dtype=torch.float32, device=w1.device)
            bfull[:hidden] = b1
            wp, bp = self.hold(split_bf16(full)), self.hold(bfull)
        wdd, bdd = self._dw_weights(wd, bd)
        self._add("uavsal_expand_dw3x3", (*x.act(), n, h, w, cin, wp.data_ptr(), kp, bp.data_ptr(), hidden, stride,
                                          wdd.data_ptr(), bdd.data_ptr(), *out.act()), tag)

    def dwproj(self, hid: Buf, n, h, w, wd, bd, w2d, bias, out: Buf, res: Optional[Buf] = None, tag=""):
        """Fused depthwise 3x3 + BN + ReLU6 -> 1x1 project + BN (+ residual) from the fp32 / q16 hidden tensor (stride 1).
        wd / w2d: W specs, or folded tensors ([9][hidden], bd) / ((cout, hidden), bias)."""
        cout, hidden = (w2d.cout, w2d.cin) if isinstance(w2d, W) else w2d.shape
        assert hid.plain and hid.c == hidden and self.engine == "tc"
        assert not hid.q16 or hidden % 128 == 0, "q16 rows feed the tensor-core kernel only"
        assert (hidden % 128 == 0 and cout % 64 == 0 and cout <= 256) or ((hidden, cout) == (32, 16) and res is None)
        if (hidden, cout) == (32, 16) and getattr(self, "dwproj32_params", True):
            # features.1: all 848 weights go into the kernel's parameter block as HOST arrays (constant-bank FFMA operands)
            if isinstance(wd, W):
                hwd, hbd = self.packed(wd, W_COLS_F32, host=True)
            else:
                hwd, hbd = wd.detach().float().cpu().contiguous(), bd.detach().float().cpu().contiguous()
                self.keep += [hwd, hbd]
            if isinstance(w2d, W):
                hw2, hb2 = self.packed(w2d, W_ROWS_F32, host=True)
            else:
                hw2, hb2 = w2d.detach().float().cpu().contiguous(), bias.detach().float().cpu().contiguous()
                self.keep += [hw2, hb2]
            self._add("uavsal_dw_project32_hw", (hid.ptr, hid.ld, n, h, w, hwd.data_ptr(), hbd.data_ptr(), hw2.data_ptr(), hb2.data_ptr(), *out.act()), tag)
            return
        wdd, bdd = self._dw_weights(wd, bd)
        if isinstance(w2d, W):
            wp, b = self.packed(w2d, W_ROWS_SPLIT, cout, hidden)
        else:
            wp, b = self.hold(pack_pw_tc(w2d, hidden)), self.hold(bias.float())
        r = res.act() if res is not None else NULL_ACT
        self._add("uavsal_dw_project", (hid.ptr, hid.ld, n, h, w, hidden, wdd.data_ptr(), bdd.data_ptr(),
                                        wp.data_ptr(), hidden, cout, b.data_ptr(),
                                        (F_RESIDUAL if res is not None else 0) | (F_HID_Q16 if hid.q16 else 0), self.terms,
                                        *r, *out.act()), tag)

    def mbconv(self, x: Buf, n, h, w, w1: W, wd: W, w2: W, out: Buf, res: Optional[Buf] = None, tag=""):
        """Whole stride-1 inverted-residual block in one kernel (mbconv.cu): expand + BN + ReLU6 -> depthwise 3x3 + BN + ReLU6 ->
        project + BN (+ residual); the hidden tensor never reaches HBM.  cin <= 64, cout % 8 == 0 and <= 64.  The kernel works on
        64-channel chunks of the hidden tensor and a project N that is a multiple of 16: other widths are passed zero-padded (padded
        hidden channels are exactly 0 through both ReLU6s and meet zero project weights; padded outputs are never stored)."""
        hid_true, cin, cout = w1.cout, w1.cin, w2.cout
        kp1 = _pad8(cin)
        hidden, cout16 = (hid_true + 63) // 64 * 64, (cout + 15) // 16 * 16
        assert self.engine == "tc" and not x.plain and x.c in (cin, kp1) and w2.cin == hid_true and wd.cout == hid_true
        assert kp1 <= 64 and cout % 8 == 0 and cout <= 64 and out.c >= cout and (res is None or not res.plain)
        w1p, b1 = self.packed(w1, W_ROWS_SPLIT, hidden, kp1)
        wdd, bdd = self.packed(wd, W_COLS_F32, hidden)
        w2p, b2 = self.packed(w2, W_ROWS_SPLIT, cout16, hidden)
        r = res.act() if res is not None else NULL_ACT
        self._add("uavsal_mbconv_fused", (*x.act(), n, h, w, kp1, w1p.data_ptr(), kp1, b1.data_ptr(), hidden, wdd.data_ptr(), bdd.data_ptr(),
                                          w2p.data_ptr(), cout, b2.data_ptr(), F_RESIDUAL if res is not None else 0, self.terms, *r, *out.act()), tag)

    def pw(self, x: Buf, m: int, w2d, bias, flags: int, out: Buf, res: Optional[Buf] = None, tag=""):
        """Pointwise conv as GEMM.  w2d: W, or folded fp32 (N, K_logical) + bias; x.c may be padded beyond K_logical."""
        n, k = (w2d.cout, w2d.cin) if isinstance(w2d, W) else w2d.shape
        kpad = _pad8(k)
        assert x.c in (k, kpad) and out.c >= n and n % 8 == 0, (x.c, k, n)
        r = res.act() if res is not None else NULL_ACT
        if res is not None:
            flags |= F_RESIDUAL
        assert not x.plain and (res is None or not res.plain), "fp32 / q16 rows are only consumed by the depthwise kernels"
        if out.f32:
            assert self.engine == "tc"
            flags |= F_OUT_F32
        elif out.q16:
            assert self.engine == "tc" and flags == F_RELU6 and res is None, "q16 rows hold a ReLU6 output"
            flags |= F_OUT_Q16
        tc = self.engine != "simt"
        if isinstance(w2d, W):
            wp, b = self.packed(w2d, W_ROWS_SPLIT if tc else W_COLS_F32, n, kpad)
            bp = b.data_ptr()
        else:
            b = self.hold(bias.float()) if bias is not None else None
            bp = b.data_ptr() if b is not None else 0
            wp = self.hold(pack_pw_tc(w2d, kpad) if tc else pack_pw_simt(w2d, kpad))
        if tc:
            self._add("uavsal_pw_gemm", (*x.act(), m, kpad, wp.data_ptr(), kpad, n, bp, flags, self.terms_arg, *r, *out.act()), tag)
        else:
            self._add("uavsal_pw_gemm_simt", (*x.act(), m, kpad, wp.data_ptr(), n, bp, flags, *r, *out.act()), tag)

    def _conv_weights(self, w4d, bias, gates: int = 1):
        """3x3 conv weights: W or a folded (Cout, Cin, 3, 3) tensor (+ bias) -> ([2][Cout][9 Cin] split | [9 Cin][Cout] fp32, bias ptr)."""
        tc = self.engine != "simt"
        if isinstance(w4d, W):
            wp, b = self.packed(w4d, W_ROWS_SPLIT if tc else W_COLS_F32, gates=gates)
            return wp, (b.data_ptr() if (w4d.bn is not None or w4d.bias is not None) else 0)
        w4 = w4d.detach().float()
        if gates > 1:
            w4 = interleave_gates(w4, w4.shape[0] // gates)
        w2d = conv3x3_as_2d(w4)
        bp = 0
        if bias is not None:
            bb = bias.detach().float()
            bp = self.hold(interleave_gates(bb, bb.shape[0] // gates) if gates > 1 else bb).data_ptr()
        return self.hold(split_bf16(w2d) if tc else w2d.t().contiguous()), bp

    def conv3x3(self, x: Buf, n, h, w, c, w4d, bias, flags, out: Buf, tag=""):
        cout = w4d.cout if isinstance(w4d, W) else w4d.shape[0]
        wp, bp = self._conv_weights(w4d, bias)
        if self.engine != "simt":
            self._add("uavsal_conv3x3", (*x.act(), n, h, w, c, wp.data_ptr(), cout, bp, flags, self.terms_arg, *out.act()), tag)
        else:
            self._add("uavsal_conv3x3_simt", (*x.act(), n, h, w, c, wp.data_ptr(), cout, bp, flags, *out.act()), tag)

    # ---- alternative backbones (ResNet / VGG) ----
    def conv_first(self, x: torch.Tensor, kind: int, n, h, w, ws: W, stride: int, flags: int, out: Buf, tag=""):
        """First conv of a ResNet (7x7 s2 + BN) / VGG (3x3 s1 + bias) from the raw frame tensor, 3 -> 64 channels."""
        assert ws.cin == 3 and ws.cout == 64 and ws.taps in (9, 49)
        wd, bd = self.packed(ws, W_COLS_F32)
        self._add("uavsal_conv_first", (x.data_ptr(), kind, n, h, w, 7 if ws.taps == 49 else 3, stride, wd.data_ptr(), bd.data_ptr(), flags, *out.act()), tag)

    def maxpool(self, x: Buf, n, h, w, c, k, stride, pad, out: Buf, tag=""):
        self._add("uavsal_maxpool", (*x.act(), n, h, w, c, k, stride, pad, *out.act()), tag)

    def add_act(self, a: Buf, b: Buf, rows, c, flags, out: Buf, tag=""):
        self._add("uavsal_add_act", (*a.act(), *b.act(), rows, c, flags, *out.act()), tag)

    def bilinear(self, x: Buf, n_src, hs, ws, c, out: Buf, n_dst, hd, wd, tag="", src_group=0, dst_group=0):
        self._add("uavsal_bilinear_ac", (*x.act(), n_src, hs, ws, c, *out.act(), n_dst, hd, wd, src_group, dst_group), tag)

    def tdiff(self, x: Buf, n, hw, c, out: Buf, tag=""):
        """``self.call_group`` (frames per reference call, 0 = the whole batch is one call) places the mirrored edges."""
        self._add("uavsal_tdiff_cat", (*x.act(), n, hw, c, *out.act(), int(getattr(self, "call_group", 0))), tag)

    def ctx_sum(self, x: Buf, b, t, hw, c, out: Buf, tag=""):
        self._add("uavsal_ctx_sum", (*x.act(), b, t, hw, c, *out.act()), tag)

    def twa(self, x: Buf, h0: Buf, t_steps, h, w, c, w4d, seq: Buf, tag="", batch: int = 1):
        """``batch`` independent sequences (x / seq: batch*t_steps frames, sequence-major; h0: batch frames) advance together.
        w4d: W or the (c, 2c, 3, 3) gate-conv weight (no bias on the UAVSal path, model.py:328)."""
        wp, _ = self._conv_weights(w4d, None)
        if self.engine != "simt":
            gx = self.tensor((batch * t_steps * h * w, c)) if (self.engine == "tc" and c % 64 == 0) else None     # hoisted W_x*x_t workspace
            sync = self.tensor((max(1, batch * ((h + 15) // 16) * ((w + 7) // 8)),), dtype=torch.int32) if gx is not None else None
            self._add("uavsal_twa_sequence", (*x.act(), *h0.act(), t_steps, h, w, c, wp.data_ptr(), 0, self.terms_arg,
                                              gx.data_ptr() if gx is not None else 0, *seq.act(), batch,
                                              sync.data_ptr() if sync is not None else 0), tag)
        else:
            self._add("uavsal_twa_sequence", (*x.act(), *h0.act(), t_steps, h, w, c, 0, wp.data_ptr(), self.terms, 0, *seq.act(), batch, 0), tag)

    def lstm(self, x: Buf, h0: Buf, c_state: torch.Tensor, b, t_steps, h, w, cin, ch, w4d, bias, seq: Buf, tag=""):
        """w4d: W (weight + optional conv bias) or the (4ch, cin+ch, 3, 3) tensor + bias; rows are re-ordered gate-interleaved."""
        wp, bp = self._conv_weights(w4d, bias, gates=4)
        if self.engine != "simt":
            args = (*x.act(), *h0.act(), c_state.data_ptr(), b, t_steps, h, w, cin, ch, wp.data_ptr(), 0, bp, self.terms_arg, *seq.act())
        else:
            args = (*x.act(), *h0.act(), c_state.data_ptr(), b, t_steps, h, w, cin, ch, 0, wp.data_ptr(), bp, self.terms, *seq.act())
        self._add("uavsal_convlstm_sequence", args, tag)

    def dw_dot_sigmoid(self, hid: Buf, n, h, w, c, wd, bd, wproj, bias, out: torch.Tensor, tag=""):
        """Readout tail fused: depthwise 3x3 + BN + ReLU6 on the fp32 / q16 hidden tensor -> 1-output project + BN + sigmoid.
        wd: W or [9][C] + bd; wproj: W (1 x C project + BN) or the folded (C,) vector + the folded scalar bias."""
        assert hid.plain and hid.c == c
        ws = self.tensor((n * h * w, (c + 63) // 64))
        wdd, bdd = self._dw_weights(wd, bd)
        wv, bias = self._dot_weights(wproj, bias)
        self._add("uavsal_dw3x3_dot_sigmoid_q16" if hid.q16 else "uavsal_dw3x3_dot_sigmoid", (hid.ptr, hid.ld, n, h, w, c, wdd.data_ptr(), bdd.data_ptr(),
                                               wv.data_ptr(), float(bias), ws.data_ptr(), out.data_ptr()), tag)

    def _dot_weights(self, wproj, bias):
        if isinstance(wproj, W):
            wv, b = self.packed(wproj, W_ROWS_F32)
            return wv, (float(b[0].item()) if self.mode != "measure" else 0.0)        # one scalar crosses to the host at plan build
        return self.hold(wproj.float()), float(bias)

    def dot_sigmoid(self, x: Buf, rows, k, wgt, bias, out: torch.Tensor, tag=""):
        wv, bias = self._dot_weights(wgt, bias)
        self._add("uavsal_dot_sigmoid", (*x.act(), rows, k, wv.data_ptr(), float(bias), out.data_ptr()), tag)

    def post_u8(self, maps: torch.Tensor, n, hs, ws, hd, wd, out_u8: torch.Tensor, tag=""):
        fm = self.tensor((n,), torch.float32)
        self._add("uavsal_post_u8", (maps.data_ptr(), n, hs, ws, hd, wd, fm.data_ptr(), out_u8.data_ptr()), tag)

    # ---- execution ----
    def mark_split(self):
        """Ops emitted after this point form the plan's "back" part (the recurrent tail of a UAVSal call, which depends on
        the previous call's state); the ops before it form the "front", which a runner may overlap with another call."""
        self.split = len(self.ops)

    def _range(self, part: Optional[str]):
        split = getattr(self, "split", None)
        if part is None or split is None:
            return 0, len(self.ops)
        return (0, split) if part == "front" else (split, len(self.ops))

    def run(self, upto: Optional[int] = None, part: Optional[str] = None):
        if self.device.type != "cuda" or self.mode == "measure":
            raise RuntimeError("uavsal-b200 kernels are CUDA (sm_100a) only; there is no CPU path")
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        lo, hi = self._range(part)
        ops = self.ops[lo:hi] if upto is None else self.ops[:upto]
        for op in ops:
            rc = op.fn(*op.args, stream)
            if rc:
                _ext.check(rc, op.name + ("[" + op.tag + "]" if op.tag else ""))

    def capture(self):
        """Warm up once eagerly, then record the op list into a CUDA graph (two graphs when the plan is split)."""
        self.run()
        torch.cuda.synchronize(self.device)
        self.graphs = {}
        parts = ["front", "back"] if getattr(self, "split", None) is not None else [None]
        for part in parts:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.run(part=part)
            self.graphs[part] = g
        self.graph = self.graphs.get(None)

    def launch(self, part: Optional[str] = None):
        graphs = getattr(self, "graphs", None)
        if graphs:
            if part is None and None not in graphs:
                graphs["front"].replay()
                graphs["back"].replay()
            else:
                graphs[part].replay()
        else:
            self.run(part=part)

    @property
    def num_launches(self) -> int:
        """Kernel launches per run (sequence ops launch one kernel per step; post_u8 launches two)."""
        n = 0
        for op in self.ops:
            if op.name == "uavsal_twa_sequence":
                n += op.args[6] * (1 if op.args[13] else op.args[17]) + (1 if op.args[13] else 0)
            elif op.name == "uavsal_convlstm_sequence":
                n += op.args[8] * (1 if self.engine != "simt" else op.args[7])
            elif op.name in ("uavsal_post_u8", "uavsal_dw3x3_dot_sigmoid", "uavsal_dw3x3_dot_sigmoid_q16"):
                n += 2
            else:
                n += 1
        return n
