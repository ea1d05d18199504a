"""Kernel micro-benchmarks on one B200 (CUDA events, L2 flushed between repetitions).  Dev tool.
    python tools/microbench.py [gemm|dw|conv|twa|all]"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from iip_uavsal_saliency_b200 import _ext
from iip_uavsal_saliency_b200.engine import Plan, pack_dw

dev = torch.device("cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(plan, reps=5):
    plan.run(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); plan.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def gemm(engine, m, k, n, res=False, terms=3, f32=False, q16=False):
    p = Plan(dev, terms, engine)
    a = p.alloc(m, k); a.t.normal_()
    o = p.alloc_q16(m, n) if q16 else p.alloc_f32(m, n) if f32 else p.alloc(m, n)
    r = p.alloc(m, n) if res else None
    w = torch.randn(n, k, device=dev) / k ** 0.5
    p.pw(a, m, w, torch.zeros(n, device=dev), 1, o, res=r)
    ms = timeit(p)
    fl = 2.0 * m * k * n
    by = 4.0 * m * (k + n * (2 if res else 1)) - (2.0 * m * n if q16 else 0)
    print("pw[%s,t%d%s] m=%d k=%d n=%d res=%d: %.1f us  %.1f TF/s(alg)  %.0f GB/s" % (engine, terms, ",q16out" if q16 else ",f32out" if f32 else "", m, k, n, res, ms * 1e3, fl / ms / 1e9, by / ms / 1e6), flush=True)


def dw(fast, n, h, w, c, stride, f32=False, q16=False, dil=1):
    _ext.load().uavsal_set_option(2, fast)
    p = Plan(dev, 3, "tc")
    x = p.alloc_q16(n * h * w, c) if q16 else p.alloc_f32(n * h * w, c) if f32 else p.alloc(n * h * w, c)
    x.t.random_(-32768, 32767) if q16 else x.t.normal_()
    ho, wo = (h, w) if stride == 1 else ((h - 1) // 2 + 1, (w - 1) // 2 + 1)
    o = p.alloc(n * ho * wo, c)
    p.dw(x, n, h, w, c, stride, dil, p.hold(pack_dw(torch.randn(c, 1, 3, 3))), p.hold(torch.zeros(c)), True, o)
    ms = timeit(p)
    by = 4.0 * n * c * (h * w + ho * wo) - (2.0 * n * c * h * w if q16 else 0)
    print("dw[fast=%d%s%s] n=%d %dx%d c=%d s=%d: %.1f us  %.0f GB/s" % (fast, ",q16in" if q16 else ",f32in" if f32 else "", ",dil=%d" % dil if dil > 1 else "", n, h, w, c, stride, ms * 1e3, by / ms / 1e6), flush=True)
    _ext.load().uavsal_set_option(2, 1)


def expdw(n, h, w, cin, hidden, stride):
    p = Plan(dev, 3, "tc")
    x = p.alloc(n * h * w, cin); x.t.normal_()
    ho, wo = (h, w) if stride == 1 else ((h - 1) // 2 + 1, (w - 1) // 2 + 1)
    o = p.alloc(n * ho * wo, hidden)
    p.expdw(x, n, h, w, torch.randn(hidden, cin, device=dev) / cin ** 0.5, torch.zeros(hidden, device=dev), stride,
            pack_dw(torch.randn(hidden, 1, 3, 3)), torch.zeros(hidden), o)
    ms = timeit(p)
    by = 4.0 * n * (h * w * cin + ho * wo * hidden)
    print("expand+dw n=%d %dx%d cin=%d hidden=%d s=%d: %.1f us  %.0f GB/s (in+out)" % (n, h, w, cin, hidden, stride, ms * 1e3, by / ms / 1e6), flush=True)


def mbblock(n, h, w, inp, oup, fused):
    """a dwBlock (stride 1) as one kernel (uavsal_mbconv_fused) or as expand GEMM -> depthwise -> project GEMM"""
    from iip_uavsal_saliency_b200 import model as M
    blk = M.dwBlock(inp, oup).eval().cuda()
    p = Plan(dev, 3, "tc")
    p.fuse_mbconv = fused
    x = p.alloc(n * h * w, inp); x.t.normal_()
    blk._emit(p, x, n, h, w)
    ms = timeit(p)
    by = 4.0 * n * h * w * (inp + oup + (inp if inp == oup else 0))
    print("dwBlock %d->%d->%d n=%d %dx%d %s: %.1f us  %.0f GB/s (block in+out)  [%s]" % (inp, 6 * inp, oup, n, h, w, "fused" if fused else "3 kernels",
          ms * 1e3, by / ms / 1e6, ", ".join(o.name.replace("uavsal_", "") for o in p.ops)), flush=True)


def dwproj(n, h, w, hidden, co, res=False, terms=3, q16=False):
    p = Plan(dev, terms, "tc")
    if q16:
        hb = p.alloc_q16(n * h * w, hidden); hb.t.random_(-32768, 32767)
    else:
        hb = p.alloc_f32(n * h * w, hidden); hb.t.uniform_(0, 6)
    o = p.alloc(n * h * w, co)
    r = p.alloc(n * h * w, co) if res else None
    p.dwproj(hb, n, h, w, pack_dw(torch.randn(hidden, 1, 3, 3)), torch.zeros(hidden), torch.randn(co, hidden, device=dev) / hidden ** 0.5,
             torch.zeros(co, device=dev), o, res=r)
    ms = timeit(p)
    print("dw+project[t%d%s] n=%d %dx%d hidden=%d co=%d: %.1f us  %.1f TF/s(alg)  %.0f GB/s(hidden read)" %
          (terms, ",q16" if q16 else "", n, h, w, hidden, co, ms * 1e3, 2.0 * n * h * w * hidden * co / ms / 1e9, (2.0 if q16 else 4.0) * n * h * w * hidden / ms / 1e6), flush=True)


def readout(n, h, w, c, q16=False):
    p = Plan(dev, 3, "tc")
    if q16:
        hb = p.alloc_q16(n * h * w, c); hb.t.random_(-32768, 32767)
    else:
        hb = p.alloc_f32(n * h * w, c); hb.t.uniform_(0, 6)
    out = p.tensor((n, 1, h, w))
    p.dw_dot_sigmoid(hb, n, h, w, c, pack_dw(torch.randn(c, 1, 3, 3)).to(dev), torch.zeros(c, device=dev), torch.randn(c, device=dev) / c ** 0.5, 0.0, out)
    ms = timeit(p)
    print("readout dw+dot[%s] n=%d %dx%d c=%d: %.1f us  %.0f GB/s(hidden read)" % ("q16" if q16 else "f32", n, h, w, c, ms * 1e3, (2.0 if q16 else 4.0) * n * h * w * c / ms / 1e6), flush=True)


def bilinear(n_src, hs, ws, c, n_dst, hd, wd, src_group=0, dst_group=0):
    p = Plan(dev, 3, "tc")
    x = p.alloc(n_src * hs * ws, c); x.t.normal_()
    o = p.alloc(n_dst * hd * wd, c)
    p.bilinear(x, n_src, hs, ws, c, o, n_dst, hd, wd, src_group=src_group, dst_group=dst_group)
    ms = timeit(p)
    print("bilinear %dx%dx%d (%d) -> %dx%d (%d): %.1f us  %.0f GB/s (written)" % (hs, ws, c, n_src, hd, wd, n_dst, ms * 1e3, 4.0 * n_dst * hd * wd * c / ms / 1e6), flush=True)


def stem(n, h, w):
    p = Plan(dev, 3, "tc")
    x = torch.randint(0, 256, (n, h, w, 3), dtype=torch.uint8, device=dev)
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    o = p.alloc_f32(n * ho * wo, 32)
    p.stem(p.hold(x), 2, n, h, w, torch.randn(3, 3, 3, 32) * 0.3, torch.zeros(32), o)
    ms = timeit(p)
    print("stem n=%d %dx%d: %.1f us  %.0f GB/s (written)" % (n, h, w, ms * 1e3, 4.0 * n * ho * wo * 32 / ms / 1e6), flush=True)


def conv(engine, n, h, w, c, co, terms=3):
    p = Plan(dev, terms, engine)
    x = p.alloc(n * h * w, c); x.t.normal_()
    o = p.alloc(n * h * w, co)
    p.conv3x3(x, n, h, w, c, torch.randn(co, c, 3, 3, device=dev) * 0.02, torch.zeros(co, device=dev), 1, o)
    ms = timeit(p)
    fl = 2.0 * n * h * w * 9 * c * co
    print("conv3x3[%s,t%d] n=%d %dx%d c=%d co=%d: %.1f us  %.1f TF/s(alg)" % (engine, terms, n, h, w, c, co, ms * 1e3, fl / ms / 1e9), flush=True)


def twa(engine, t, h, w, c, terms=3, batch=1):
    p = Plan(dev, terms, engine)
    x = p.alloc(batch * t * h * w, c); x.t.normal_()
    h0 = p.alloc(batch * h * w, c)
    seq = p.alloc(batch * t * h * w, c)
    p.twa(x, h0, t, h, w, c, torch.randn(c, 2 * c, 3, 3, device=dev) * 0.01, seq, batch=batch)
    ms = timeit(p)
    print("twa[%s] batch=%d t=%d %dx%d c=%d: %.1f us total, %.1f us/step (incl. the hoisted x half), %.1f TF/s(alg)" % (engine, batch, t, h, w, c, ms * 1e3, ms * 1e3 / t, 2.0 * batch * t * h * w * 9 * 2 * c * c / ms / 1e9), flush=True)


def lstm(b, t, h, w, c, terms=3):
    """BASELINE config #3: ConvLSTM((h,w), c, c, (3,3), 1, batch_first=True, bias=False), x (b,t,c,h,w), h0 = c0 = 0."""
    p = Plan(dev, terms, "tc")
    x = p.alloc(b * t * h * w, c); x.t.normal_()
    h0 = p.alloc(b * h * w, c)
    cst = p.tensor((b, h * w, c))
    seq = p.alloc(b * t * h * w, c)
    wgt = torch.empty(4 * c, 2 * c, 3, 3, device=dev)
    torch.nn.init.xavier_uniform_(wgt)
    p.lstm(x, h0, cst, b, t, h, w, c, c, wgt, None, seq)
    p.run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    cst.zero_()
    e0.record(); p.run(); e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    fl = 2.0 * b * h * w * 4 * c * 9 * 2 * c * t
    print("convlstm[tc,t%d] b=%d t=%d %dx%d c=%d: %.2f ms total, %.1f us/step, %.1f TF/s(alg), %.1f TF/s(issued)" %
          (terms, b, t, h, w, c, ms, ms * 1e3 / t, fl / ms / 1e9, terms * fl / ms / 1e9), flush=True)


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    if what == "dwproj_q16":
        dwproj(120, 45, 80, 1536, 256, res=True, q16=True); dwproj(120, 45, 80, 1920, 256, q16=True); dwproj(120, 45, 80, 1152, 64, q16=True)
        dwproj(120, 45, 80, 1536, 256, res=True)
        return
    if what == "aspp":          # dilated depthwise convs of the ASPP branches (120 frames of 12x20 x 1920 channels)
        for d in (6, 12, 18):
            dw(0, 120, 12, 20, 1920, 1, dil=d); dw(2, 120, 12, 20, 1920, 1, dil=d); dw(2, 120, 12, 20, 1920, 1, q16=True, dil=d)
        gemm("tc", 28800, 320, 1920); gemm("tc", 28800, 320, 1920, q16=True)
        return
    if what == "stem":
        stem(120, 360, 640); dwproj(120, 180, 320, 32, 16)
        return
    if what == "bilinear":      # the five upsample / broadcast launches of a 120-frame plan
        bilinear(120, 12, 20, 256, 120, 45, 80); bilinear(120, 23, 40, 128, 120, 45, 80)
        bilinear(1, 45, 80, 64, 120, 45, 80); bilinear(24, 12, 20, 64, 120, 45, 80, src_group=4, dst_group=20)
        return
    if what == "glueprof":      # one launch each of the kernels reworked late in round 2, for an ncu --set full capture
        readout(120, 45, 80, 1536, q16=True); stem(120, 360, 640); bilinear(120, 12, 20, 256, 120, 45, 80)
        dw(2, 120, 12, 20, 1920, 1, q16=True, dil=6); mbblock(120, 90, 160, 24, 24, True)
        return
    if what == "pairq16prof":   # the plan's top kernel (256 -> 1536 expand conv, q16 rows out) for an ncu --set full capture
        gemm("tc", 120 * 3600, 256, 1536, q16=True)
        return
    if what == "q16prof":       # one launch each for ncu (after a warm-up launch)
        dwproj(120, 45, 80, 1536, 256, res=True, q16=True); readout(120, 45, 80, 1536, q16=True)
        return
    if what == "q16":           # fp32 rows vs 16-bit fixed-point rows for the widest hidden tensors (120 frames of 45x80)
        M = 120 * 3600
        for q in (False, True):
            gemm("tc", M, 256, 1536, f32=not q, q16=q); gemm("tc", M, 320, 1920, f32=not q, q16=q); gemm("tc", M, 192, 1152, f32=not q, q16=q)
            dwproj(120, 45, 80, 1536, 256, res=True, q16=q); dwproj(120, 45, 80, 1920, 256, q16=q); dwproj(120, 45, 80, 1152, 64, q16=q)
            readout(120, 45, 80, 1536, q16=q)
            dw(2, 24, 45, 80, 1536, 2, f32=not q, q16=q)
        return
    M = 72000
    if what in ("gemm", "all"):
        for eng in ("tc1", "tc"):
            gemm(eng, M, 256, 1536); gemm(eng, M, 1536, 256, res=True); gemm(eng, M, 320, 1920); gemm(eng, M, 1920, 256)
            gemm(eng, M, 256, 256); gemm(eng, M, 256, 32); gemm(eng, 20 * 180 * 320, 16, 96); gemm(eng, 20 * 90 * 160, 144, 24, res=True)
            gemm(eng, 4800, 320, 1920); gemm(eng, M, 256, 1536, terms=1)
    if what in ("dw", "all"):
        for fast in (0, 2):
            dw(fast, 20, 45, 80, 1536, 1); dw(fast, 20, 180, 320, 96, 2); dw(fast, 20, 180, 320, 32, 1); dw(fast, 20, 90, 160, 144, 1); dw(fast, 20, 23, 40, 384, 1)
    if what == "dwbigf32":
        dw(2, 20, 45, 80, 1536, 1, True)
    if what == "dwbig":
        dw(int(sys.argv[2]) if len(sys.argv) > 2 else 2, 20, 45, 80, 1536, 1)
    if what == "r2":
        gemm("tc", M, 256, 1536); gemm("tc", M, 256, 1536, f32=True); gemm("tc", M, 1536, 256, res=True); gemm("tc", M, 320, 1920, f32=True)
        gemm("tc", M, 1920, 256); gemm("tc", M, 256, 256); gemm("tc", M, 192, 1152, f32=True); gemm("tc", M, 1152, 64)
        gemm("tc", M, 256, 1536, terms=1, f32=True); gemm("tc", 20 * 180 * 320, 16, 96, f32=True); gemm("tc", 20 * 90 * 160, 144, 24, res=True)
        conv("tc", 20, 45, 80, 448, 256); twa("tc", 20, 45, 80, 256)
        for f in (False, True):
            dw(2, 20, 45, 80, 1536, 1, f); dw(2, 20, 180, 320, 96, 2, f); dw(2, 20, 180, 320, 32, 1, f); dw(2, 20, 90, 160, 144, 1, f); dw(2, 20, 23, 40, 384, 1, f)
    if what == "expdw":
        expdw(20, 180, 320, 16, 96, 2); expdw(20, 90, 160, 24, 144, 1); expdw(20, 90, 160, 24, 144, 2); expdw(20, 45, 80, 32, 192, 1)
        expdw(20, 45, 80, 32, 192, 2)
    if what == "res":
      lib = _ext.load()
      for mask in (0, 1 << 20):
        lib.uavsal_set_option(3, mask)
        print("--- residual loads:", "row-strided" if mask else "coalesced")
        gemm("tc", 432000, 32, 256, res=True); gemm("tc", 432000, 256, 256, res=True); gemm("tc", 110400, 384, 64, res=True)
        gemm("tc", 1728000, 144, 24, res=True); gemm("tc", M, 1536, 256, res=True)
      lib.uavsal_set_option(3, 0)
    if what == "f32set":       # the expand GEMMs of a 120-frame plan (fp32 hidden rows out)
        for (m, k, n) in ((432000, 256, 1536), (432000, 320, 1920), (432000, 192, 1152), (432000, 32, 192), (432000, 64, 384), (1728000, 24, 144),
                          (110400, 64, 384), (110400, 96, 576), (28800, 160, 960)):
            gemm("tc", m, k, n, f32=True)
    if what == "smallk":       # HBM-bound expand GEMMs: the epilogue's store rate is what matters
        gemm("tc", 1728000, 24, 144, f32=True); gemm("tc", 432000, 32, 192, f32=True); gemm("tc", 432000, 32, 256, res=True)
    if what == "mbconv_trace":
        lib = _ext.load()
        lib.uavsal_set_option(3, 1 << 22)
        from iip_uavsal_saliency_b200 import model as M
        blk = M.dwBlock(64, 32).eval().cuda()
        p = Plan(dev, 3, "tc")
        x = p.alloc(120 * 45 * 80, 64); x.t.normal_()
        blk._emit(p, x, 120, 45, 80)
        p.run(); torch.cuda.synchronize()
        lib.uavsal_set_option(3, 0)
    if what == "mbconv":
        for fused in (False, True):
            mbblock(120, 45, 80, 64, 32, fused); mbblock(120, 45, 80, 32, 32, fused); mbblock(120, 23, 40, 64, 64, fused)
            mbblock(120, 90, 160, 24, 24, fused)
    if what == "pairbig":
        gemm("tc", 432000, 256, 1536, f32=True)
    if what == "wr":
        gemm("tc", 432000, 32, 256, res=True); gemm("tc", 432000, 32, 192, f32=True); gemm("tc", 432000, 64, 384, f32=True)
    if what == "expdw2":
        expdw(20, 180, 320, 16, 96, 2)
    if what == "expdw1":
        expdw(20, 45, 80, 64, 384, 1)
    if what == "lstm":
        lstm(8, 64, 45, 80, 256); lstm(8, 64, 45, 80, 256, terms=1)
    if what == "dwproj":
        dwproj(20, 45, 80, 1536, 256); dwproj(20, 45, 80, 1536, 256, res=True); dwproj(20, 45, 80, 1920, 256); dwproj(20, 45, 80, 1152, 64)
        dwproj(20, 45, 80, 1536, 256, terms=1); dw(2, 20, 45, 80, 1536, 1, True); gemm("tc", M, 1536, 256, res=True)
    if what == "dwproj32only":
        dwproj(20, 180, 320, 32, 16)
    if what == "dwproj32":
        dwproj(20, 180, 320, 32, 16); dw(2, 20, 180, 320, 32, 1, True); gemm("tc", 20 * 180 * 320, 32, 16)
    if what == "dwproj1":
        dwproj(20, 45, 80, 1536, 256)
    if what == "twa2":
        lib = _ext.load()
        for mode in (0, 1):
            lib.uavsal_set_option(7, mode)
            print("--- twa step kernel mode", mode)
            twa("tc", 20, 45, 80, 256); twa("tc", 60, 45, 80, 256)
        lib.uavsal_set_option(7, 1)
    if what == "twa3":
        lib = _ext.load()
        for mode in (1, 3):
            lib.uavsal_set_option(7, mode)
            print("--- twa mode", mode, "(1 = one launch per step, 3 = one launch per sequence)")
            for batch in (1, 2, 4):
                twa("tc", 60, 45, 80, 256, batch=batch)
        lib.uavsal_set_option(7, 1)
    if what == "twa_trace":
        lib = _ext.load()
        lib.uavsal_set_option(3, 1 << 22)
        lib.uavsal_set_option(7, 3)
        p = Plan(dev, 3, "tc")
        t, h, w, c, batch = 12, 45, 80, 256, 2
        x = p.alloc(batch * t * h * w, c); x.t.normal_()
        h0 = p.alloc(batch * h * w, c); seq = p.alloc(batch * t * h * w, c)
        p.twa(x, h0, t, h, w, c, torch.randn(c, 2 * c, 3, 3, device=dev) * 0.01, seq, batch=batch)
        p.run(); torch.cuda.synchronize()
        lib.uavsal_set_option(3, 0)
        lib.uavsal_set_option(7, 1)
    if what == "twa_ablate":
        lib = _ext.load()
        for mask, name in ((0, "full"), (1 << 16, "no-mma"), (1 << 17, "no-epilogue-io"), (1 << 18, "no-B"), (7 << 16, "none of them")):
            lib.uavsal_set_option(3, mask)
            print("---", name)
            twa("tc", 60, 45, 80, 256)
        lib.uavsal_set_option(3, 0)
    if what == "twa_bn":
        lib = _ext.load()
        for bn in (64, 128):
            lib.uavsal_set_option(8, bn)
            print("--- twa step kernel bn", bn)
            twa("tc", 60, 45, 80, 256)
            lib.uavsal_set_option(3, 1 << 16); twa("tc", 60, 45, 80, 256); lib.uavsal_set_option(3, 0)
        lib.uavsal_set_option(8, 64)
    if what == "twa_t1":
        lib = _ext.load()
        for terms in (3, 1):
            print("--- terms", terms)
            twa("tc", 60, 45, 80, 256, terms)
            lib.uavsal_set_option(3, 1 << 16); twa("tc", 60, 45, 80, 256, terms); lib.uavsal_set_option(3, 0)
    if what == "stages":
        lib = _ext.load()
        for st in (1, 2, 3, 6):
            lib.uavsal_set_option(5, st)
            print("--- max stages", st)
            gemm("tc", M, 1536, 256, res=True); gemm("tc", M, 1536, 256, terms=1); gemm("tc", M, 256, 1536, f32=True); conv("tc", 20, 45, 80, 448, 256)
        lib.uavsal_set_option(5, 8)
    if what == "cluster":
        lib = _ext.load()
        for cl in (1, 2):
            lib.uavsal_set_option(4, cl)
            print("--- cluster", cl)
            gemm("tc", M, 256, 1536); gemm("tc", M, 1536, 256, res=True); gemm("tc", M, 320, 1920); gemm("tc", M, 1920, 256)
            gemm("tc", M, 256, 256); gemm("tc", M, 192, 1152); gemm("tc", M, 1152, 64); gemm("tc", M, 256, 1536, terms=1)
            gemm("tc", 20 * 180 * 320, 16, 96); gemm("tc", 20 * 90 * 160, 144, 24, res=True)
            conv("tc", 20, 45, 80, 448, 256)
    if what == "ablate":
        # timing ablations of the persistent GEMM: bit16 no MMA, bit17 no global stores, bit18 no B loads, bit19 no A loads
        lib = _ext.load()
        lib.uavsal_set_option(4, int(sys.argv[2]) if len(sys.argv) > 2 else 2)
        for mask, name in ((0, "full"), (1 << 16, "no-mma"), (1 << 17, "no-store"), (1 << 18, "no-B"), (1 << 19, "no-A"),
                           (3 << 18, "no-A no-B"), (3 << 16, "no-mma no-store"), (7 << 17, "no-store no-A no-B"), (0xF << 16, "nothing")):
            lib.uavsal_set_option(3, mask)
            print("---", name)
            gemm("tc", M, 256, 1536, f32=True); gemm("tc", M, 1536, 256, res=True); gemm("tc", M, 256, 1536, terms=1, f32=True)
        lib.uavsal_set_option(3, 0)
    if what == "gemmbig":
        gemm("tc", M, 256, 1536); gemm("tc", M, 1536, 256, res=True)
    if what in ("conv", "all"):
        for eng in ("tc1", "tc"):
            conv(eng, 20, 45, 80, 448, 256); conv(eng, 20, 45, 80, 448, 256, terms=1)
    if what in ("twa", "all"):
        for eng in ("tc1", "tc"):
            twa(eng, 20, 45, 80, 256)


if __name__ == "__main__":
    main()
