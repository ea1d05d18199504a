"""Plan builder + executor: turns the module tree into a flat list of C-ABI kernel calls over a
pre-allocated activation arena (NHWC, split-bf16 planes — see include/uavsal_b200.h).

Nothing here computes on the CPU: a plan can be *built* on any device (the CPU test-suite checks shapes,
packing and the op list that way) but ``Plan.run`` requires CUDA tensors and the sm_100a library.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import torch

from . import _ext

F_RELU6, F_RESIDUAL, F_SIGMOID, F_OUT_F32 = 1, 2, 4, 8
PLANE_F32 = -1            # UAVSAL_PLANE_F32: the activation is plain fp32 rows, not split-bf16 planes
BN_EPS = 1e-5


def _pad8(c: int) -> int:
    return (c + 7) // 8 * 8


def out_size(n: int, stride: int) -> int:
    """3x3, padding = dilation, stride 1|2 (model.py:67-69)."""
    return n if stride == 1 else (n - 1) // 2 + 1


class Buf:
    """A (rows x c) activation living in a (2, rows, ld) bf16 tensor at channel offset ``off`` — or, when ``f32``, in a
    plain (rows, ld) fp32 tensor (the hidden tensor between an expand conv and its depthwise conv)."""

    __slots__ = ("t", "rows", "c", "ld", "off", "f32")

    def __init__(self, t: torch.Tensor, rows: int, c: int, ld: int, off: int = 0, f32: bool = False):
        self.t, self.rows, self.c, self.ld, self.off, self.f32 = t, rows, c, ld, off, f32

    @property
    def ptr(self) -> int:
        return self.t.data_ptr() + (4 if self.f32 else 2) * self.off

    @property
    def plane(self) -> int:
        return PLANE_F32 if self.f32 else self.rows * self.ld

    def act(self) -> Tuple[int, int, int]:
        return (self.ptr, self.plane, self.ld)

    def slot(self, off: int, c: int) -> "Buf":
        assert off % 8 == 0 and off + c <= self.ld
        return Buf(self.t, self.rows, c, self.ld, self.off + off, self.f32)

    def to_float(self) -> torch.Tensor:
        """fp32 (rows, c) reconstruction hi + lo (debug / tests)."""
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
    def __init__(self, device: torch.device, terms: int = 3, engine: str = "tc"):
        assert terms in (1, 3) and engine in ("tc", "tc1", "simt")
        self.device = torch.device(device)
        self.terms = terms
        self.engine = engine
        self.ops: List[Op] = []
        self.keep: List[torch.Tensor] = []       # packed weights / scratch kept alive with the plan
        self.named: Dict[str, object] = {}       # name -> Buf / tensor (inputs, outputs, debug taps)
        self.graph = None
        self.arena_bytes = 0

    # ---- allocation ----
    def alloc(self, rows: int, c: int, ld: Optional[int] = None) -> Buf:
        ld = _pad8(c) if ld is None else ld
        t = torch.zeros((2, rows, ld), dtype=torch.bfloat16, device=self.device)
        self.arena_bytes += t.numel() * 2
        self.keep.append(t)
        return Buf(t, rows, c, ld)

    def alloc_f32(self, rows: int, c: int) -> Buf:
        """fp32 rows (only the persistent tcgen05 GEMM / the stem write them, only the TMA depthwise kernel reads them)."""
        ld = _pad8(c)
        t = torch.zeros((rows, ld), dtype=torch.float32, device=self.device)
        self.arena_bytes += t.numel() * 4
        self.keep.append(t)
        return Buf(t, rows, c, ld, 0, True)

    @property
    def f32_hidden(self) -> bool:
        return self.engine == "tc"

    def tensor(self, shape, dtype=torch.float32) -> torch.Tensor:
        t = torch.zeros(shape, dtype=dtype, device=self.device)
        self.arena_bytes += t.numel() * t.element_size()
        self.keep.append(t)
        return t

    def hold(self, t: torch.Tensor) -> torch.Tensor:
        t = t.to(self.device).contiguous()
        self.keep.append(t)
        return t

    def _add(self, name: str, args: Sequence, tag: str = ""):
        fn = getattr(_ext.load(), name) if self.device.type == "cuda" else None
        self.ops.append(Op(name, fn, tuple(args), tag))

    # ---- ops ----
    def pack_nchw(self, src: torch.Tensor, n, c, h, w, dst: Buf, tag=""):
        self._add("uavsal_pack_nchw_f32", (src.data_ptr(), n, c, h, w, *dst.act(), _pad8(c)), tag)

    def unpack_nchw(self, src: Buf, n, c, h, w, dst: torch.Tensor, tag=""):
        self._add("uavsal_unpack_nchw_f32", (*src.act(), n, c, h, w, dst.data_ptr()), tag)

    def stem(self, x: torch.Tensor, kind: int, n, h, w, wgt, bias, out: Buf, tag=""):
        """wgt: folded fp32 [3][3][3][32] (ky, kx, cin, cout), bias [32].  The tcgen05 engines hand the 3.5 KB of weights over as
        HOST arrays: they travel in the kernel parameters and every FFMA reads its weight from the constant bank."""
        if self.engine != "simt" and tuple(wgt.shape) == (3, 3, 3, 32):
            wh = wgt.detach().float().cpu().contiguous()
            bh = bias.detach().float().cpu().contiguous()
            self.keep += [wh, bh]
            self._add("uavsal_stem_conv3x3s2_hw", (x.data_ptr(), kind, n, h, w, wh.data_ptr(), bh.data_ptr(), *out.act()), tag)
            return
        wd, bd = self.hold(wgt), self.hold(bias)
        self._add("uavsal_stem_conv3x3s2", (x.data_ptr(), kind, n, h, w, wd.data_ptr(), bd.data_ptr(), *out.act()), tag)

    def dw(self, x: Buf, n, h, w, c, stride, dil, wgt, bias, relu6, out: Buf, tag=""):
        self._add("uavsal_dw3x3", (*x.act(), n, h, w, c, stride, dil, wgt.data_ptr(), bias.data_ptr(), int(relu6), *out.act()), tag)

    def expdw(self, x: Buf, n, h, w, w1: torch.Tensor, b1: torch.Tensor, stride: int, wd: torch.Tensor, bd: torch.Tensor, out: Buf, tag=""):
        """Fused 1x1 expand + BN + ReLU6 -> depthwise 3x3 + BN + ReLU6 (cin <= 32).  w1: folded fp32 (hidden, cin_logical)."""
        hidden, k = w1.shape
        cin = x.c
        assert not x.f32 and cin % 8 == 0 and k <= cin <= 32 and hidden % 8 == 0
        kp = (cin + 15) // 16 * 16
        hp = (hidden + 63) // 64 * 64
        full = torch.zeros((hp, kp), dtype=torch.float32, device=w1.device)
        full[:hidden, :k] = w1
        bfull = torch.zeros((hp,), dtype=torch.float32, device=w1.device)
        bfull[:hidden] = b1
        wp, bp = self.hold(split_bf16(full)), self.hold(bfull)
        self._add("uavsal_expand_dw3x3", (*x.act(), n, h, w, cin, wp.data_ptr(), kp, bp.data_ptr(), hidden, stride,
                                          self.hold(wd).data_ptr(), self.hold(bd.float()).data_ptr(), *out.act()), tag)

    def dwproj(self, hid: Buf, n, h, w, wd: torch.Tensor, bd: torch.Tensor, w2d: torch.Tensor, bias: torch.Tensor, out: Buf,
               res: Optional[Buf] = None, tag=""):
        """Fused depthwise 3x3 + BN + ReLU6 -> 1x1 project + BN (+ residual) from the fp32 hidden tensor (stride 1)."""
        cout, hidden = w2d.shape
        assert hid.f32 and hid.c == hidden and self.engine == "tc"
        assert (hidden % 128 == 0 and cout % 64 == 0 and cout <= 256) or ((hidden, cout) == (32, 16) and res is None)
        if (hidden, cout) == (32, 16) and getattr(self, "dwproj32_params", True):
            # features.1: all 848 weights go into the kernel's parameter block as HOST arrays (constant-bank FFMA operands)
            host = [t.detach().float().cpu().contiguous() for t in (wd, bd, w2d, bias)]
            self.keep += host
            self._add("uavsal_dw_project32_hw", (hid.ptr, hid.ld, n, h, w, *[t.data_ptr() for t in host], *out.act()), tag)
            return
        wp = self.hold(pack_pw_tc(w2d, hidden))
        b = self.hold(bias.float())
        r = res.act() if res is not None else NULL_ACT
        self._add("uavsal_dw_project", (hid.ptr, hid.ld, n, h, w, hidden, self.hold(wd).data_ptr(), self.hold(bd.float()).data_ptr(),
                                        wp.data_ptr(), hidden, cout, b.data_ptr(), F_RESIDUAL if res is not None else 0, self.terms,
                                        *r, *out.act()), tag)

    def pw(self, x: Buf, m: int, w2d: torch.Tensor, bias: Optional[torch.Tensor], flags: int, out: Buf,
           res: Optional[Buf] = None, tag=""):
        """Pointwise conv as GEMM.  w2d: folded fp32 (N, K_logical); x.c may be padded beyond K_logical."""
        n, k = w2d.shape
        kpad = _pad8(k)
        assert x.c in (k, kpad) and out.c >= n and n % 8 == 0, (x.c, k, n)
        b = self.hold(bias.float()) if bias is not None else None
        bp = b.data_ptr() if b is not None else 0
        r = res.act() if res is not None else NULL_ACT
        if res is not None:
            flags |= F_RESIDUAL
        assert not x.f32 and (res is None or not res.f32), "fp32 rows are only consumed by the depthwise kernel"
        if out.f32:
            assert self.engine == "tc"
            flags |= F_OUT_F32
        if self.engine != "simt":
            wp = self.hold(pack_pw_tc(w2d, kpad))
            self._add("uavsal_pw_gemm", (*x.act(), m, kpad, wp.data_ptr(), kpad, n, bp, flags, self.terms, *r, *out.act()), tag)
        else:
            wp = self.hold(pack_pw_simt(w2d, kpad))
            self._add("uavsal_pw_gemm_simt", (*x.act(), m, kpad, wp.data_ptr(), n, bp, flags, *r, *out.act()), tag)

    def conv3x3(self, x: Buf, n, h, w, c, w4d: torch.Tensor, bias, flags, out: Buf, tag=""):
        cout = w4d.shape[0]
        w2d = conv3x3_as_2d(w4d)
        b = self.hold(bias.float()) if bias is not None else None
        bp = b.data_ptr() if b is not None else 0
        if self.engine != "simt":
            wp = self.hold(split_bf16(w2d))
            self._add("uavsal_conv3x3", (*x.act(), n, h, w, c, wp.data_ptr(), cout, bp, flags, self.terms, *out.act()), tag)
        else:
            wp = self.hold(w2d.t().contiguous())
            self._add("uavsal_conv3x3_simt", (*x.act(), n, h, w, c, wp.data_ptr(), cout, bp, flags, *out.act()), tag)

    def bilinear(self, x: Buf, n_src, hs, ws, c, out: Buf, n_dst, hd, wd, tag="", src_group=0, dst_group=0):
        self._add("uavsal_bilinear_ac", (*x.act(), n_src, hs, ws, c, *out.act(), n_dst, hd, wd, src_group, dst_group), tag)

    def tdiff(self, x: Buf, n, hw, c, out: Buf, tag=""):
        """``self.call_group`` (frames per reference call, 0 = the whole batch is one call) places the mirrored edges."""
        self._add("uavsal_tdiff_cat", (*x.act(), n, hw, c, *out.act(), int(getattr(self, "call_group", 0))), tag)

    def ctx_sum(self, x: Buf, b, t, hw, c, out: Buf, tag=""):
        self._add("uavsal_ctx_sum", (*x.act(), b, t, hw, c, *out.act()), tag)

    def twa(self, x: Buf, h0: Buf, t_steps, h, w, c, w4d: torch.Tensor, seq: Buf, tag="", batch: int = 1):
        """``batch`` independent sequences (x / seq: batch*t_steps frames, sequence-major; h0: batch frames) advance together."""
        w2d = conv3x3_as_2d(w4d.detach().float())
        if self.engine != "simt":
            wp = self.hold(split_bf16(w2d))
            gx = self.tensor((batch * t_steps * h * w, c)) if (self.engine == "tc" and c % 64 == 0) else None     # hoisted W_x*x_t workspace
            self._add("uavsal_twa_sequence", (*x.act(), *h0.act(), t_steps, h, w, c, wp.data_ptr(), 0, self.terms,
                                              gx.data_ptr() if gx is not None else 0, *seq.act(), batch), tag)
        else:
            wp = self.hold(w2d.t().contiguous())
            self._add("uavsal_twa_sequence", (*x.act(), *h0.act(), t_steps, h, w, c, 0, wp.data_ptr(), self.terms, 0, *seq.act(), batch), tag)

    def lstm(self, x: Buf, h0: Buf, c_state: torch.Tensor, b, t_steps, h, w, cin, ch, w4d, bias, seq: Buf, tag=""):
        wi = interleave_gates(w4d.detach().float(), ch)
        w2d = conv3x3_as_2d(wi)
        bp = 0
        if bias is not None:
            bb = self.hold(interleave_gates(bias.detach().float(), ch))
            bp = bb.data_ptr()
        if self.engine != "simt":
            wp = self.hold(split_bf16(w2d))
            args = (*x.act(), *h0.act(), c_state.data_ptr(), b, t_steps, h, w, cin, ch, wp.data_ptr(), 0, bp, self.terms, *seq.act())
        else:
            wp = self.hold(w2d.t().contiguous())
            args = (*x.act(), *h0.act(), c_state.data_ptr(), b, t_steps, h, w, cin, ch, 0, wp.data_ptr(), bp, self.terms, *seq.act())
        self._add("uavsal_convlstm_sequence", args, tag)

    def dw_dot_sigmoid(self, hid: Buf, n, h, w, c, wd: torch.Tensor, bd: torch.Tensor, wproj: torch.Tensor, bias: float, out: torch.Tensor, tag=""):
        """Readout tail fused: depthwise 3x3 + BN + ReLU6 on the fp32 hidden tensor -> 1-output project + BN + sigmoid."""
        assert hid.f32 and hid.c == c
        ws = self.tensor((n * h * w, (c + 63) // 64))
        self._add("uavsal_dw3x3_dot_sigmoid", (hid.ptr, hid.ld, n, h, w, c, self.hold(wd).data_ptr(), self.hold(bd.float()).data_ptr(),
                                               self.hold(wproj.float()).data_ptr(), float(bias), ws.data_ptr(), out.data_ptr()), tag)

    def dot_sigmoid(self, x: Buf, rows, k, wgt: torch.Tensor, bias: float, out: torch.Tensor, tag=""):
        wv = self.hold(wgt.float())
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
        if self.device.type != "cuda":
            raise RuntimeError("uavsal-b200 kernels are CUDA (sm_100a) only; there is no CPU path")
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _ext.load().uavsal_set_option(1, 1 if self.engine == "tc1" else 2)      # tcgen05 kernel generation (process-global)
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
            elif op.name in ("uavsal_post_u8", "uavsal_dw3x3_dot_sigmoid"):
                n += 2
            else:
                n += 1
        return n
