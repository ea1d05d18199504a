"""CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional re-statement of what /root/reference computes on the UAVSal inference path, written against
a plain ``state_dict`` (no nn.Module classes of the product are used) with torch CPU fp32 primitives as the
arithmetic (conv2d / batch_norm / hardtanh / interpolate are the same third-party ops the reference itself
calls).  Every function cites the reference lines it follows.

Parity pin: tests/test_oracle_golden.py compares these functions with outputs of the real reference
(oracle/make_golden.py, run in the authoring container where /root/reference exists).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

EPS = 2.2204e-16          # utils_score_torch.py:13, utils_data.py:7
BN_EPS = 1e-5             # torch.nn.BatchNorm2d default used by model.py:70,95

# torchvision mobilenet_v2 inverted_residual_setting (t, c, n, s) — the architecture ReMobileNetV2 wraps
# (model_feature.py:59-60); block i lives at features.{i}
_MBV2_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
                 (6, 320, 1, 1)]


# ---------------------------------------------------------------------------------------------------
# building blocks
# ---------------------------------------------------------------------------------------------------
def _bn(sd, p, x):
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        False, 0.0, BN_EPS)


def basic_conv(sd, p, x, stride=1, dilation=1, groups=1):
    """BasicConv2d = conv(no bias) + BN + ReLU6 (model.py:65-72); torchvision Conv2dNormActivation likewise."""
    w = sd[p + ".0.weight"]
    k = w.shape[-1]
    pad = dilation * (k - 1) // 2
    x = F.conv2d(x, w, None, stride, pad, dilation, groups)
    return F.hardtanh(_bn(sd, p + ".1", x), 0.0, 6.0)


def dw_block(sd, p, x, stride=1, dilation=1, res_connect=None):
    """dwBlock / InvertedResidual with expand (model.py:74-103): pw+BN+ReLU6 → dw3x3+BN+ReLU6 → pw+BN (+x)."""
    inp = x.shape[1]
    h = basic_conv(sd, p + ".conv.0", x)
    h = basic_conv(sd, p + ".conv.1", h, stride=stride, dilation=dilation, groups=h.shape[1])
    h = F.conv2d(h, sd[p + ".conv.2.weight"])
    h = _bn(sd, p + ".conv.3", h)
    use_res = stride == 1 and inp == h.shape[1]
    if res_connect is not None:
        use_res = bool(res_connect) and use_res
    return x + h if use_res else h


def mobilenet_v2_features(sd, p, x) -> Tuple[torch.Tensor, ...]:
    """ReMobileNetV2.forward (model_feature.py:62-69): features[0:2],[2:4],[4:7],[7:14],[14:18]."""
    x = basic_conv(sd, p + ".0", x, stride=2)
    # features.1: expand_ratio 1 → dw+BN+ReLU6, pw+BN, no residual (16 != 32)
    h = basic_conv(sd, p + ".1.conv.0", x, groups=x.shape[1])
    x = _bn(sd, p + ".1.conv.2", F.conv2d(h, sd[p + ".1.conv.1.weight"]))
    outs = {1: x}
    idx = 2
    for t, c, n, s in _MBV2_SETTING[1:]:
        for i in range(n):
            x = dw_block(sd, "%s.%d" % (p, idx), x, stride=s if i == 0 else 1)
            outs[idx] = x
            idx += 1
    return outs[1], outs[3], outs[6], outs[13], outs[17]


_RESNET_CFG = {"resnet18": ("basic", (2, 2, 2, 2)), "resnet34": ("basic", (3, 4, 6, 3)), "resnet50": ("bottleneck", (3, 4, 6, 3)),
               "resnet101": ("bottleneck", (3, 4, 23, 3)), "resnet152": ("bottleneck", (3, 8, 36, 3))}
_VGG16_CFG = (64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M")


def _conv_bn(sd, pc, pb, x, stride=1, relu=True):
    w = sd[pc + ".weight"]
    x = _bn(sd, pb, F.conv2d(x, w, None, stride, w.shape[-1] // 2))
    return F.relu(x) if relu else x


def resnet_features(sd, p, x, name="resnet50"):
    """ReResNet.forward (model_feature.py:92-103) over torchvision's ResNet (resnet.py: BasicBlock / Bottleneck with the stride on
    the 3x3 conv, downsample = conv1x1(stride) + BN): returns x0 (after the max pool), layer1 .. layer4."""
    kind, blocks = _RESNET_CFG[name]
    x = _conv_bn(sd, p + ".conv1", p + ".bn1", x, stride=2)
    x = F.max_pool2d(x, 3, 2, 1)
    outs = [x]
    for li, nb in enumerate(blocks):
        for bi in range(nb):
            q = "%s.layer%d.%d" % (p, li + 1, bi)
            stride = 2 if (bi == 0 and li > 0) else 1
            idn = x
            if q + ".downsample.0.weight" in sd:
                idn = _conv_bn(sd, q + ".downsample.0", q + ".downsample.1", x, stride=stride, relu=False)
            if kind == "basic":
                y = _conv_bn(sd, q + ".conv1", q + ".bn1", x, stride=stride)
                y = _conv_bn(sd, q + ".conv2", q + ".bn2", y, relu=False)
            else:
                y = _conv_bn(sd, q + ".conv1", q + ".bn1", x)
                y = _conv_bn(sd, q + ".conv2", q + ".bn2", y, stride=stride)
                y = _conv_bn(sd, q + ".conv3", q + ".bn3", y, relu=False)
            x = F.relu(y + idn)
        outs.append(x)
    return tuple(outs)


def vgg_features(sd, p, x):
    """ReVGG.forward (model_feature.py:117-128) over torchvision's vgg16.features.  The reference finds the pooling layers by
    zipping ``features.modules()`` (whose first element is the Sequential itself) with range(100), so its split points sit one
    PAST every max pool: each of the five levels ends with its pooling layer (64 @ 1/2, 128 @ 1/4, 256 @ 1/8, 512 @ 1/16, 512 @ 1/32)."""
    outs = []
    idx = 0
    for v in _VGG16_CFG:
        if v == "M":
            x = F.max_pool2d(x, 2, 2)
            outs.append(x)
            idx += 1
        else:
            x = F.relu(F.conv2d(x, sd["%s.%d.weight" % (p, idx)], sd["%s.%d.bias" % (p, idx)], 1, 1))
            idx += 2
    return tuple(outs)


def backbone_features(sd, p, x, cnn_type="mobilenet_v2"):
    if cnn_type == "mobilenet_v2":
        return mobilenet_v2_features(sd, p + ".features", x)
    if cnn_type == "vgg16":
        return vgg_features(sd, p + ".features", x)
    return resnet_features(sd, p, x, cnn_type)


def srfnet(sd, p, x, trace=None, cnn_type="mobilenet_v2"):
    """uavsal_srfnet_aspp.forward (model.py:139-158)."""
    _, _, c3, c4, c5 = backbone_features(sd, p + ".features", x, cnn_type)
    if trace is not None:
        trace.update(c3=c3, c4=c4, c5=c5)
    a1 = basic_conv(sd, p + ".lv5_aspp1", c5)
    a2 = dw_block(sd, p + ".lv5_aspp2", c5, dilation=6)
    a3 = dw_block(sd, p + ".lv5_aspp3", c5, dilation=12)
    a4 = dw_block(sd, p + ".lv5_aspp4", c5, dilation=18)
    x5 = basic_conv(sd, p + ".conv_lv5", torch.cat((a1, a2, a3, a4), 1))
    x4 = basic_conv(sd, p + ".conv_lv4", c4)
    x3 = basic_conv(sd, p + ".conv_lv3", c3)
    size = c3.shape[2:]
    x5 = F.interpolate(x5, size=size, mode="bilinear", align_corners=True)
    x4 = F.interpolate(x4, size=size, mode="bilinear", align_corners=True)
    cat = torch.cat((x5, x4, x3), 1)
    if trace is not None:
        trace.update(aspp_cat=torch.cat((a1, a2, a3, a4), 1), sf_cat=cat)
    return basic_conv(sd, p + ".conv_last", cat)


def te_conv_sub(sd, p, x):
    """teConv_sub.forward (model.py:188-208): neighbour differences over the WHOLE call batch (quirk Q2)."""
    x1 = basic_conv(sd, p + ".reduce_conv", x)
    n = x1.shape[0]
    prev = torch.cat([x1[1:2], x1[:-1]], 0)      # frame 0 pairs with frame 1 (model.py:194)
    nxt = torch.cat([x1[1:], x1[-2:-1]], 0)      # last frame pairs with n-2 (model.py:198)
    d = torch.cat([x1 - prev, x1 - nxt], 1)
    d[0] = torch.cat([x1[1] - x1[0], x1[0] - x1[1]], 0)            # model.py:194 (sign of first half flips)
    d[n - 1] = torch.cat([x1[-1] - x1[-2], x1[-2] - x1[-1]], 0)    # model.py:198 (sign of second half flips)
    h = dw_block(sd, p + ".sub_conv", d, res_connect=False)
    return basic_conv(sd, p + ".last_conv", h)


def st_block(sd, p, x):
    """STBlock.forward, fu_type='sum', res_connect=True (model.py:235-249)."""
    sp = dw_block(sd, p + ".stconv_sp.spconv", x, res_connect=False)
    te = te_conv_sub(sd, p + ".stconv_te", x)
    return x + basic_conv(sd, p + ".stconv_last", sp + te)


def twa_sequence(w, x_seq, h):
    """ConvTWA.forward, one layer, batch 1 (model_convlstm.py:333-383) with ConvTWACell.forward (276-292).
    x_seq (T,C,H,W), h (1,C,H,W) → (T,C,H,W), h_last."""
    outs = []
    for t in range(x_seq.shape[0]):
        xt = x_seq[t:t + 1]
        i = torch.sigmoid(F.conv2d(torch.cat([xt, h], 1), w, None, 1, 1))
        h = i * xt + (1 - i) * h
        outs.append(h)
    return torch.cat(outs, 0), h


def lstm_sequence(w, b, x, h, c):
    """ConvLSTM.forward, one layer, batch_first (model_convlstm.py:168-218) with ConvLSTMCell.forward
    (111-126): gate order i, f, o, g.  x (B,T,C,H,W) → (B,T,Ch,H,W), (h, c)."""
    ch = h.shape[1]
    outs = []
    for t in range(x.shape[1]):
        cc = F.conv2d(torch.cat([x[:, t], h], 1), w, b, 1, 1)
        ci, cf, co, cg = torch.split(cc, ch, dim=1)
        i, f, o, g = torch.sigmoid(ci), torch.sigmoid(cf), torch.sigmoid(co), torch.tanh(cg)
        c = f * c + i * g
        h = o * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1), (h, c)


def uavsal_forward(sd: Dict[str, torch.Tensor], x, cb, h0, time_dims=5, num_stblock=2,
                   bias_type=(1, 1, 1), trace: Optional[dict] = None, cnn_type: str = "mobilenet_v2"):
    """UAVSal.forward (model.py:341-375).  x (N,3,H,W) normalised fp32, cb=[gauss (N,8,h,w), ob (N,20,h,w)],
    h0 (1,256,h,w).  Returns out (N,1,h,w), h_last (1,256,h,w).  ``trace`` collects named intermediates.
    With h0 = (h, c) the recurrence is the ConvLSTM of the UAVSAL_LSTM ablation (model.py:960-1076) and (h, c) is returned."""
    tr = trace if trace is not None else {}
    with torch.no_grad():
        x = srfnet(sd, "sfnet", x, tr, cnn_type)
        tr["sfnet"] = x
        for i in range(num_stblock):
            x = st_block(sd, "st_layer.%d" % i, x)
            tr["st_layer.%d" % i] = x
        x = dw_block(sd, "fust_layer.0", x)
        tr["fust"] = x
        if any(bias_type):
            fu = []
            if bias_type[0]:
                g = dw_block(sd, "gauss_cb_layer.1", dw_block(sd, "gauss_cb_layer.0", cb[0]))
                fu.append(g)
                tr["cb_gauss"] = g
            if bias_type[1]:
                o = dw_block(sd, "ob_cb_layer.1", dw_block(sd, "ob_cb_layer.0", cb[1]))
                fu.append(o)
                tr["cb_ob"] = o
            if bias_type[2]:
                n, c, hh, ww = x.shape
                b = n // time_dims
                s = x.contiguous().view(b, time_dims, c, hh, ww).sum(1)              # model.py:357-358
                s = dw_block(sd, "cxt_cb_prior.1", dw_block(sd, "cxt_cb_prior.0", s, stride=2), stride=2)
                s = F.interpolate(s, size=(hh, ww), mode="bilinear", align_corners=True)
                s = s.repeat(time_dims, 1, 1, 1)                                       # model.py:361 (quirk Q3)
                fu.append(s)
                tr["cb_cxt"] = s
            xcb = dw_block(sd, "fucb_layer.0", torch.cat(fu, 1))
            tr["fucb"] = xcb
            x = dw_block(sd, "fucbst_layer.0", torch.cat([x, xcb], 1))
            tr["fucbst"] = x
        if isinstance(h0, (list, tuple)):
            # UAVSAL_LSTM.forward (model.py:1065-1068): 4-gate ConvLSTM over the call's frames, batch 1; h0 = (h, c)
            seq5, (h, c) = lstm_sequence(sd["rnn.cell_list.0.rnn_conv.weight"], None, x[None], h0[0], h0[1])
            seq, h = seq5[0], (h, c)
        else:
            seq, h = twa_sequence(sd["rnn.cell_list.0.rnn_conv.weight"], x, h0)
        tr["rnn"] = seq
        out = torch.sigmoid(dw_block(sd, "conv_out_st", seq))
        tr["out"] = out
    return out, h


# ---------------------------------------------------------------------------------------------------
# data helpers (utils_data.py)
# ---------------------------------------------------------------------------------------------------
_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def normalize_data(u8: np.ndarray) -> np.ndarray:
    """normalize_data for uint8 (N,3,H,W) (utils_data.py:43-65): float32 /255, then (x-mean)/std with the
    python-float mean/std applied to a float32 array (numpy keeps float32)."""
    ims = u8.astype(np.float32) / 255.0
    for c in range(3):
        ims[:, c] = (ims[:, c] - _MEAN[c]) / _STD[c]
    return ims


def st_get_gaussmaps(height, width, nb):
    """utils_data.py:391-412."""
    e = height / width
    e1 = (1 - e) / 2
    e2 = e1 + e
    sig = e * np.arange(1, nb + 1) / 16
    xt = np.repeat(np.linspace(0.0, 1.0, width)[None, :, None], height, 0).repeat(nb, 2)
    yt = np.repeat(np.linspace(e1, e2, height)[:, None, None], width, 1).repeat(nb, 2)
    return 1 / (2 * np.pi * sig * sig + EPS) * np.exp(
        -((xt - 0.5) ** 2 / (2 * sig ** 2 + EPS) + (yt - 0.5) ** 2 / (2 * sig ** 2 + EPS)))


def gauss_priors(height=45, width=80, nb=8) -> np.ndarray:
    """get_guasspriors when the .mat is absent (utils_data.py:453-457): per-channel min-max normalised."""
    g = st_get_gaussmaps(height, width, nb)
    g = (g - g.min((0, 1))) / (g.max((0, 1)) - g.min((0, 1)) + EPS)
    return g.astype(np.float32)


def resize_linear_f32(src: np.ndarray, dst_h: int, dst_w: int) -> np.ndarray:
    """cv2.resize(float32, INTER_LINEAR) restated (OpenCV imgproc resize.cpp, unpinned dependency —
    SURVEY §8(c)): half-pixel centres, source index clamped, float weights, horizontal pass then vertical."""
    sh, sw = src.shape

    def taps(dst, srcn):
        scale = srcn / dst
        f = (np.arange(dst) + 0.5) * scale - 0.5
        s = np.floor(f).astype(np.int64)
        w = (f - s).astype(np.float32)
        lo = s < 0
        s[lo] = 0
        w[lo] = 0.0
        hi = s >= srcn - 1
        s[hi] = srcn - 1
        w[hi] = 0.0
        s1 = np.minimum(s + 1, srcn - 1)
        return s, s1, w

    x0, x1, wx = taps(dst_w, sw)
    y0, y1, wy = taps(dst_h, sh)
    src = src.astype(np.float32)
    rows = src[:, x0] * (np.float32(1) - wx) + src[:, x1] * wx
    return rows[y0] * (np.float32(1) - wy)[:, None] + rows[y1] * wy[:, None]


def postprocess_predictions(pred: np.ndarray, shape_r: int, shape_c: int) -> np.ndarray:
    """utils_data.py:289-303: letterbox-inverse resize, then /max*255 (float)."""
    pr, pc = pred.shape
    if shape_r / pr > shape_c / pc:
        new_cols = (pc * shape_r) // pr
        big = resize_linear_f32(pred, shape_r, new_cols)
        o = (new_cols - shape_c) // 2
        img = big[:, o:o + shape_c]
    else:
        new_rows = (pr * shape_c) // pc
        big = resize_linear_f32(pred, new_rows, shape_c)
        o = (new_rows - shape_r) // 2
        img = big[o:o + shape_r, :]
    return img / np.max(img) * 255


def im2uint8(img: np.ndarray) -> np.ndarray:
    """utils_data.py:68-75 (np2mat with dtype uint8, :77-82): clip, round-half-even, cast."""
    if img.dtype == np.uint8:
        return img
    return np.rint(np.clip(img, 0, 255)).astype(np.uint8)


def demo_test_clip(sd, frames_u8: np.ndarray, gauss: np.ndarray, ob: np.ndarray, time_dims=5, batch_size=4,
                   out_hw=None, bias_type=(1, 1, 1)):
    """The per-video loop of Demo_Test.test (Demo_Test.py:68-91) on an in-memory clip.
    frames_u8 (F,H,W,3) RGB; gauss (h,w,8) / ob (h,w,20) float32 prior maps (already at map size).
    Returns float maps (F',1,h,w) and uint8 (F',H,W) with F' = floor(F/time_dims)*time_dims (quirk Q1)."""
    F_, H, W, _ = frames_u8.shape
    h, w = gauss.shape[:2]
    out_hw = out_hw or (H, W)
    count_bs = F_ // time_dims
    keep = count_bs * time_dims
    vid = frames_u8[:keep].transpose(0, 3, 1, 2)
    per_call = batch_size * time_dims
    state = torch.zeros(1, 256, h, w)
    maps, u8 = [], []
    for i in range(math.ceil(count_bs / batch_size)):
        chunk = vid[i * per_call:(i + 1) * per_call]
        n = chunk.shape[0]
        x = torch.from_numpy(normalize_data(chunk))
        cb = [torch.from_numpy(np.repeat(gauss.transpose(2, 0, 1)[None], n, 0).copy()),
              torch.from_numpy(np.repeat(ob.transpose(2, 0, 1)[None], n, 0).copy())]
        out, state = uavsal_forward(sd, x, cb, state, time_dims=time_dims, bias_type=bias_type)
        o = out.numpy()
        maps.append(o)
        for j in range(n):
            u8.append(im2uint8(postprocess_predictions(o[j, 0], out_hw[0], out_hw[1])))
    return np.concatenate(maps, 0), np.stack(u8, 0)


# ---------------------------------------------------------------------------------------------------
# metrics (utils_score_torch.py:20-50, 180-218) — fp32, reductions over (H,W) per (n, channel)
# ---------------------------------------------------------------------------------------------------
def _sum(x):
    return torch.sum(x, (2, 3), keepdim=True)


def _mean(x):
    return torch.mean(x, (2, 3), keepdim=True)


def _std(x):
    return torch.std(x, (2, 3), keepdim=True)       # unbiased, utils_score_torch.py:49


def _amax(x):
    return torch.amax(x, (2, 3), keepdim=True)


def _amin(x):
    return torch.amin(x, (2, 3), keepdim=True)


def metric_kl(y_pred, y_true):
    """utils_score_torch.py:180-185."""
    t = y_true[:, 0:1]
    t = t / (_sum(t) + EPS)
    p = y_pred / (_sum(y_pred) + EPS)
    return torch.sum(t * torch.log(t / (p + EPS) + EPS), (2, 3))


def metric_cc(y_pred, y_true):
    """utils_score_torch.py:188-197."""
    t = y_true[:, 0:1]
    t = (t - _mean(t)) / (_std(t) + EPS)
    p = (y_pred - _mean(y_pred)) / (_std(y_pred) + EPS)
    t = t - _mean(t)
    p = p - _mean(p)
    r1 = torch.sum(t * p, (2, 3))
    r2 = torch.sqrt(torch.sum(p * p, (2, 3)) * torch.sum(t * t, (2, 3)))
    return r1 / (r2 + EPS)


def metric_nss(y_pred, y_true):
    """utils_score_torch.py:200-204."""
    f = y_true[:, 1:2]
    p = (y_pred - _mean(y_pred)) / (_std(y_pred) + EPS)
    return torch.sum(f * p, (2, 3)) / (torch.sum(f, (2, 3)) + EPS)


def metric_sim(y_pred, y_true):
    """utils_score_torch.py:207-218."""
    t = y_true[:, 0:1]
    t = (t - _amin(t)) / (_amax(t) - _amin(t) + EPS)
    p = (y_pred - _amin(y_pred)) / (_amax(y_pred) - _amin(y_pred) + EPS)
    t = t / (_sum(t) + EPS)
    p = p / (_sum(p) + EPS)
    return torch.sum(torch.min(t, p), (2, 3))


def metrics4(y_pred, y_true):
    """(N,4) columns CC, NSS, KLD, SIM."""
    return torch.cat([metric_cc(y_pred, y_true), metric_nss(y_pred, y_true), metric_kl(y_pred, y_true),
                      metric_sim(y_pred, y_true)], 1)


# ---------------------------------------------------------------------------------------------------
# AUC metrics (utils_score_torch.py:53-177).  The reference draws from the GLOBAL torch / numpy generators (jitter :82,
# random pixels :103 / :143): seed them before calling for reproducible values; the draws below happen in the same order.
# ---------------------------------------------------------------------------------------------------
def _minmax_norm(y_pred):
    return (y_pred - _amin(y_pred)) / (_amax(y_pred) - _amin(y_pred) + EPS)


def auc_j(S, F):
    """utils_score_torch.py:53-74.  S (P,) fp32 in [0,1], F (P,) bool."""
    if not torch.any(S > 0) or not torch.any(F > 0):
        return torch.tensor(float("nan"))
    S_fix = S[F]
    n_fix, n_pixels = S_fix.shape[0], S.shape[0]
    thresholds, _ = torch.sort(S_fix, descending=True)
    tp, fp = torch.zeros(n_fix + 2), torch.zeros(n_fix + 2)
    tp[-1] = 1
    fp[-1] = 1
    tp[1:-1] = (torch.arange(0, n_fix) + 1) / float(n_fix)
    # above_th[i] = #{S >= thresholds[i]}: the reference loops over the thresholds; one sort + searchsorted gives the same counts
    s_sorted, _ = torch.sort(S)
    above_th = n_pixels - torch.searchsorted(s_sorted, thresholds.contiguous(), right=False)
    fp[1:-1] = (above_th - torch.arange(0, n_fix) - 1) / float(n_pixels - n_fix)
    return torch.trapz(tp, fp)


def metric_auc_j(y_pred, y_true, jitter=1):
    """utils_score_torch.py:77-88."""
    f = y_true[:, 1:2] > 0.5
    if jitter == True:  # noqa: E712  (the reference's comparison)
        y_pred = y_pred + (torch.rand(y_pred.shape) * 1e-7).to(y_pred.device)
    y_pred = _minmax_norm(y_pred)
    S, F = torch.flatten(y_pred, 1, -1), torch.flatten(f, 1, -1)
    return torch.Tensor([auc_j(S[i], F[i]) for i in range(S.shape[0])]).unsqueeze(1)


_trapz = getattr(np, "trapezoid", None) or np.trapz          # np.trapz (the reference's call) was renamed in numpy 2


def _auc_sampled(S_fix, S_rand, n_den, step_size=0.1):
    """Shared tail of auc_b / auc_s (utils_score_torch.py:106-119, 146-158): S_rand (k, n_rep)."""
    n_fix, n_rep = len(S_fix), S_rand.shape[1]
    auc = np.zeros(n_rep) * np.nan
    for rep in range(n_rep):
        thresholds = np.r_[0:np.max(np.r_[S_fix, S_rand[:, rep]]):step_size][::-1]
        tp, fp = np.zeros(len(thresholds) + 2), np.zeros(len(thresholds) + 2)
        tp[-1] = 1
        fp[-1] = 1
        for k, thresh in enumerate(thresholds):
            tp[k + 1] = np.sum(S_fix >= thresh) / float(n_fix)
            fp[k + 1] = np.sum(S_rand[:, rep] >= thresh) / float(n_den)
        auc[rep] = _trapz(tp, fp)
    return np.mean(auc)


def auc_b(S, F, n_rep=100):
    """utils_score_torch.py:91-120 (numpy; S float32 (P,), F bool (P,))."""
    if not np.any(S > 0) or not np.any(F > 0):
        return torch.tensor(float("nan"))
    S_fix = S[F]
    n_fix, n_pixels = len(S_fix), len(S)
    r = np.random.randint(0, n_pixels, [n_fix, n_rep])
    return _auc_sampled(S_fix, S[r], n_fix)


def auc_s(S, F, Oth, n_rep=100):
    """utils_score_torch.py:135-159."""
    if not np.any(S > 0) or not np.any(F > 0):
        return torch.tensor(float("nan"))
    S_fix = S[F]
    n_fix = len(S_fix)
    ind = np.nonzero(Oth)[0]
    n_ind = len(ind)
    n_fix_oth = min(n_fix, n_ind)
    r = np.random.randint(0, n_ind, [n_ind, n_rep])[:n_fix_oth, :]
    return _auc_sampled(S_fix, S[ind[r]], n_fix_oth)


def metric_auc_b(y_pred, y_true):
    """utils_score_torch.py:123-132."""
    f = y_true[:, 1:2] > 0.5
    y_pred = _minmax_norm(y_pred)
    S, F = torch.flatten(y_pred, 1, -1).numpy(), torch.flatten(f, 1, -1).numpy()
    return torch.Tensor([auc_b(S[i], F[i]) for i in range(S.shape[0])]).unsqueeze(1)


def metric_auc_s(y_pred, y_true, shuff_map):
    """utils_score_torch.py:162-172."""
    f = y_true[:, 1:2] > 0.5
    y_pred = _minmax_norm(y_pred)
    S, F = torch.flatten(y_pred, 1, -1).numpy(), torch.flatten(f, 1, -1).numpy()
    O = torch.flatten(shuff_map, 1, -1).numpy()
    return torch.Tensor([auc_s(S[i], F[i], O[i]) for i in range(S.shape[0])]).unsqueeze(1)


# ---------------------------------------------------------------------------------------------------
# video front-end after decode: utils_data.padding (:321-343) inside preprocess_videos (:255-287).
# cv2.resize(uint8, INTER_LINEAR) is third-party code (opencv-python, unpinned by the reference; 4.13 in this image): its
# published algorithm is restated here and checked against cv2 itself in tests/test_oracle_golden.py.
# ---------------------------------------------------------------------------------------------------
def _cv_taps(dst, src, zero_at_border):
    """OpenCV resize.cpp: f = (float)((d + 0.5) * scale - 0.5), s = floor(f), 11-bit fixed-point weights rounded half to even
    (saturate_cast<short>); the x taps are clamped with a zeroed fraction at the image border, the y taps clamp their rows."""
    scale = src / dst
    d = np.arange(dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s0 = np.floor(f).astype(np.int64)
    f = (f - s0.astype(np.float32)).astype(np.float32)
    if zero_at_border:
        lo = s0 < 0
        f[lo] = 0
        s0[lo] = 0
        hi = s0 >= src - 1
        f[hi] = 0
        s0[hi] = src - 1
    a0 = np.rint((np.float32(1.0) - f) * np.float32(2048.0)).astype(np.int64)
    a1 = np.rint(f * np.float32(2048.0)).astype(np.int64)
    return s0, a0, a1


def cv_resize_linear_u8(img, dw, dh):
    """cv2.resize(img, (dw, dh)) for uint8 (H,W[,C]) images, default INTER_LINEAR: identity copy, the exact-2x case that
    OpenCV routes to its INTER_AREA 2x2 average, else horizontal pass in 11-bit fixed point and the vertical pass
    ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2 >> 2 (VResizeLinear<uchar>)."""
    sh, sw = img.shape[:2]
    a = img.astype(np.int64).reshape(sh, sw, -1)
    if (dw, dh) == (sw, sh):
        return img.copy()
    if sw == 2 * dw and sh == 2 * dh:
        out = (a[0::2, 0::2] + a[0::2, 1::2] + a[1::2, 0::2] + a[1::2, 1::2] + 2) >> 2
        return out.astype(np.uint8).reshape((dh, dw) + img.shape[2:])
    sx, ax0, ax1 = _cv_taps(dw, sw, True)
    sy, ay0, ay1 = _cv_taps(dh, sh, False)
    hp = a[:, sx] * ax0[None, :, None] + a[:, np.minimum(sx + 1, sw - 1)] * ax1[None, :, None]
    s0, s1 = hp[np.clip(sy, 0, sh - 1)], hp[np.clip(sy + 1, 0, sh - 1)]
    out = (((ay0[:, None, None] * (s0 >> 4)) >> 16) + ((ay1[:, None, None] * (s1 >> 4)) >> 16) + 2) >> 2
    return np.clip(out, 0, 255).astype(np.uint8).reshape((dh, dw) + img.shape[2:])


def letterbox_geometry(sh, sw, shape_r, shape_c):
    """utils_data.padding (:321-343): (resized width, resized height, x offset, y offset) of the aspect-preserving resize."""
    if sh / shape_r > sw / shape_c:
        new_cols = (sw * shape_r) // sh
        return new_cols, shape_r, (shape_c - min(new_cols, shape_c)) // 2, 0
    new_rows = (sh * shape_c) // sw
    return shape_c, new_rows, 0, (shape_r - min(new_rows, shape_r)) // 2


def padding(img, shape_r=480, shape_c=640, channels=3):
    """utils_data.py:321-343 with the resize restated."""
    out = np.zeros((shape_r, shape_c, channels) if channels != 1 else (shape_r, shape_c), np.uint8)
    nw, nh, ox, oy = letterbox_geometry(img.shape[0], img.shape[1], shape_r, shape_c)
    out[oy:oy + nh, ox:ox + nw] = cv_resize_linear_u8(img, nw, nh)
    return out


def preprocess_frames(frames_bgr, shape_r, shape_c, mode="RGB"):
    """preprocess_videos (:255-287) after the decode, normalize=False: (n,h,w,3) uint8 BGR frames (cv2.VideoCapture order) ->
    (n,shape_r,shape_c,3) uint8, channels swapped to RGB for mode 'RGB' (:270)."""
    ims = np.stack([padding(f, shape_r, shape_c, 3) for f in frames_bgr], 0)
    if mode == "RGB":
        ims = ims[:, :, :, [2, 1, 0]]
    elif mode != "BGR":
        raise ValueError
    return ims
