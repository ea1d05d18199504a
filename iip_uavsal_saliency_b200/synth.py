"""Seeded synthetic inputs and weights for benchmarks and parity tests (re-exported as ``oracle.synth``).

Everything here is pure numpy (``RandomState``) so the same seed gives bit-identical arrays in the
authoring container and on the GPU box; nothing reads /root/reference.

  make_clip            uint8 (F,H,W,3) RGB frames, smooth low-frequency field drifting over time + noise
                       (what ``utils_data.preprocess_videos`` would hand to Demo_Test.py:65)
  make_priors          (gauss (N,8,h,w), ob (N,20,h,w)) float32 NCHW, the layout Demo_Test.get_bias builds
  make_state_dict      a full 685-key UAVSal state dict from tests/golden/state_dict_keys.json
                       kind="stock":  the reference's init *rules*  → degenerate out≈0.5
                       kind="lively": fan-in scaled weights + non-trivial BN statistics → every layer active
  make_metric_pairs    correlated (pred, true) uint8-valued maps for CC/NSS/KLD/SIM
"""
from __future__ import annotations

import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
KEYS_JSON = os.path.join(os.path.dirname(_HERE), "tests", "golden", "state_dict_keys.json")


def _upsample_linear(a: np.ndarray, H: int, W: int) -> np.ndarray:
    """Separable linear interpolation (align-corners) of a (..., h, w) array to (..., H, W)."""
    h, w = a.shape[-2:]
    ys = np.linspace(0, h - 1, H)
    xs = np.linspace(0, w - 1, W)
    y0 = np.clip(np.floor(ys).astype(int), 0, h - 2)
    x0 = np.clip(np.floor(xs).astype(int), 0, w - 2)
    wy = (ys - y0)[:, None]
    wx = (xs - x0)[None, :]
    top = a[..., y0, :] * (1 - wy) + a[..., y0 + 1, :] * wy
    return top[..., :, x0] * (1 - wx) + top[..., :, x0 + 1] * wx


def make_clip(seed: int, frames: int, H: int = 360, W: int = 640) -> np.ndarray:
    rs = np.random.RandomState(1000 + seed)
    field = rs.rand(3, 12, 20)
    out = np.empty((frames, H, W, 3), np.uint8)
    for f in range(frames):
        field = np.clip(field + rs.randn(3, 12, 20) * 0.15 * 0.3, 0.0, 1.0)
        img = _upsample_linear(field, H, W) * 255.0
        img = img + rs.randint(-8, 9, size=(3, H, W))
        out[f] = np.clip(np.rint(img), 0, 255).astype(np.uint8).transpose(1, 2, 0)
    return out


def make_priors(n: int, h: int = 45, w: int = 80, seed: int = 0):
    """Smooth synthetic priors in [0,1] with the reference's shapes (used when the .mat files are absent
    or the map size is not 45x80, where the reference's own loader yields zeros — SURVEY Q4)."""
    rs = np.random.RandomState(2000 + seed)
    g = _upsample_linear(rs.rand(8, 5, 8), h, w).astype(np.float32)
    o = _upsample_linear(rs.rand(20, 5, 8), h, w).astype(np.float32)
    g = np.ascontiguousarray(np.broadcast_to(g[None], (n, 8, h, w)))
    o = np.ascontiguousarray(np.broadcast_to(o[None], (n, 20, h, w)))
    return g, o


def load_key_table():
    with open(KEYS_JSON) as fh:
        return json.load(fh)


def _init_rule(key: str) -> str:
    """Which of the reference's init rules applies to a conv weight (model.py:133,167,186,
    233,297,306,315,319-324,333-335; model_convlstm.py:274)."""
    if key.startswith(("gauss_cb_layer", "ob_cb_layer", "cxt_cb_prior")):
        return "kaiming_fan_in"
    if key.startswith(("fucb_layer", "fucbst_layer")):
        return "default_uniform"
    return "kaiming_fan_out"


def make_state_dict(kind: str = "lively", seed: int = 0, as_torch: bool = True):
    rs = np.random.RandomState(3000 + seed)
    table = load_key_table()
    # convs that feed a ReLU6 (gain sqrt2) vs linear convs (gain 1): a conv is linear when it is the
    # project conv of a block (key ends ".conv.2.weight", or ".conv.1.weight" in the t=1 block) or the
    # recurrent gate conv.
    sd = {}
    for key, shape, dtype in table:
        if dtype == "torch.int64":
            sd[key] = np.zeros(shape, np.int64)
            continue
        leaf = key.rsplit(".", 1)[1]
        if len(shape) == 4:
            cout, cin_g, kh, kw = shape
            fan_in = cin_g * kh * kw
            fan_out = cout * kh * kw
            if kind == "stock":
                rule = _init_rule(key)
                if rule == "kaiming_fan_out":
                    w = rs.randn(*shape) * np.sqrt(2.0 / fan_out)
                elif rule == "kaiming_fan_in":
                    w = rs.randn(*shape) * np.sqrt(2.0 / fan_in)
                else:
                    bound = 1.0 / np.sqrt(fan_in)
                    w = rs.uniform(-bound, bound, size=shape)
            else:
                linear = (key.endswith(".conv.2.weight") or key.endswith("features.1.conv.1.weight")
                          or "rnn_conv" in key)
                gain = 1.0 if linear else np.sqrt(2.0)
                w = rs.randn(*shape) * (gain / np.sqrt(fan_in))
            sd[key] = w.astype(np.float32)
        else:
            n = shape[0]
            if kind == "stock":
                val = {"weight": np.ones(n), "bias": np.zeros(n), "running_mean": np.zeros(n),
                       "running_var": np.ones(n)}[leaf]
            else:
                val = {"weight": rs.uniform(0.8, 1.2, n), "bias": rs.randn(n) * 0.1,
                       "running_mean": rs.randn(n) * 0.1, "running_var": rs.uniform(0.8, 1.2, n)}[leaf]
            sd[key] = val.astype(np.float32)
    if as_torch:
        import torch
        sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    return sd


def key_table_of(module):
    """[(key, shape, dtype)] of a module's state dict, in order (the table make_state_dict_like consumes)."""
    return [(k, list(v.shape), str(v.dtype)) for k, v in module.state_dict().items()]


def make_state_dict_like(table, seed: int = 0, as_torch: bool = True):
    """'Lively' weights for ANY key table (used for the ResNet / VGG backbone variants, whose state dicts are too large to commit):
    fan-in scaled conv weights (gain sqrt2 in front of a ReLU, 1 for the linear convs that close a residual block), non-trivial BN
    statistics, small conv biases.  Depends only on the table's order, shapes and the seed."""
    rs = np.random.RandomState(6000 + seed)
    sd = {}
    for key, shape, dtype in table:
        if "int64" in dtype:
            sd[key] = np.zeros(shape, np.int64)
            continue
        leaf = key.rsplit(".", 1)[1]
        if len(shape) == 4:
            cout, cin_g, kh, kw = shape
            linear = (key.endswith((".conv.2.weight", ".conv3.weight", ".downsample.0.weight")) or key.endswith("features.1.conv.1.weight")
                      or "rnn_conv" in key)
            gain = 1.0 if linear else np.sqrt(2.0)
            sd[key] = (rs.randn(*shape) * (gain / np.sqrt(cin_g * kh * kw))).astype(np.float32)
        else:
            n = shape[0] if shape else 1
            val = {"weight": rs.uniform(0.8, 1.2, n), "bias": rs.randn(n) * 0.1,
                   "running_mean": rs.randn(n) * 0.1, "running_var": rs.uniform(0.8, 1.2, n)}[leaf]
            sd[key] = val.astype(np.float32).reshape(shape)
    if as_torch:
        import torch
        sd = {k: torch.from_numpy(v) for k, v in sd.items()}
    return sd


def make_lstm_weight(hidden: int = 256, inp: int = 256, seed: int = 0, bias: bool = False):
    """xavier_uniform (model_convlstm.py:109) for ConvLSTMCell.rnn_conv (4*hidden, inp+hidden, 3, 3)."""
    rs = np.random.RandomState(4000 + seed)
    fan_in = (inp + hidden) * 9
    fan_out = 4 * hidden * 9
    bound = np.sqrt(6.0 / (fan_in + fan_out))
    w = rs.uniform(-bound, bound, size=(4 * hidden, inp + hidden, 3, 3)).astype(np.float32)
    b = (rs.randn(4 * hidden) * 0.1).astype(np.float32) if bias else None
    return w, b


def make_metric_pairs(n: int, H: int = 360, W: int = 640, seed: int = 0):
    """Correlated saliency pairs: density = sum of 3-8 Gaussian blobs (uint8-valued), pred = perturbed
    density (uint8-valued), fixpts = 20-200 samples from the density.  Returns float32
    pred (n,1,H,W) and true (n,2,H,W) exactly as evalscores_vid_torch feeds metric_* (utils_score_torch.py
    :538-549)."""
    rs = np.random.RandomState(5000 + seed)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    pred = np.empty((n, 1, H, W), np.float32)
    true = np.empty((n, 2, H, W), np.float32)
    for i in range(n):
        k = rs.randint(3, 9)
        dens = np.zeros((H, W), np.float32)
        pr = np.zeros((H, W), np.float32)
        for _ in range(k):
            cy, cx = rs.uniform(0.1, 0.9) * H, rs.uniform(0.1, 0.9) * W
            s = rs.uniform(0.03, 0.12) * W
            a = rs.uniform(0.3, 1.0)
            dens += a * np.exp(-((yy - cy) ** 2 + (xx - cx) ** 2) / (2 * s * s))
            jy, jx = cy + rs.randn() * 0.03 * H, cx + rs.randn() * 0.03 * W
            s2 = s * rs.uniform(0.8, 1.4)
            pr += a * rs.uniform(0.6, 1.2) * np.exp(-((yy - jy) ** 2 + (xx - jx) ** 2) / (2 * s2 * s2))
        pr += 0.05 * _upsample_linear(rs.rand(9, 16), H, W).astype(np.float32)
        dens_u8 = np.rint(dens / dens.max() * 255.0)
        pred_u8 = np.rint(pr / pr.max() * 255.0)
        nfix = rs.randint(20, 201)
        p = dens.ravel().astype(np.float64)
        p /= p.sum()
        idx = rs.choice(H * W, size=nfix, replace=False, p=p)
        fix = np.zeros(H * W, np.float32)
        fix[idx] = 1.0
        pred[i, 0] = pred_u8
        true[i, 0] = dens_u8
        true[i, 1] = fix.reshape(H, W)
    return pred, true


def make_auc_case(seed: int = 0, n: int = 6, H: int = 360, W: int = 640):
    """Inputs of the AUC metrics: `n` correlated pairs (make_metric_pairs) followed by two degenerate ones (an all-zero
    prediction; a pair without fixations - both score NaN, utils_score_torch.py:54,92,136), and a shuffle map per pair
    (fixations of the OTHER pairs, as getshufmap builds it)."""
    pred, true = make_metric_pairs(n + 2, H, W, seed=seed)
    pred[n] = 0.0
    true[n + 1, 1] = 0.0
    fix = true[:, 1]
    shuf = np.stack([np.clip(fix.sum(0) - fix[i], 0, None) for i in range(n + 2)], 0)[:, None].astype(np.float32)
    return pred, true, shuf


def make_state_dict_lstm(seed: int = 0):
    """State dict of the UAVSAL_LSTM ablation (model.py:960-1076): the UAVSal set with the recurrent conv replaced by a
    4-gate ConvLSTM kernel (4*256, 512, 3, 3), fan-in scaled so that the gates are neither saturated nor idle."""
    import torch
    sd = make_state_dict("lively", seed)
    rs = np.random.RandomState(6000 + seed)
    sd["rnn.cell_list.0.rnn_conv.weight"] = torch.from_numpy((rs.randn(1024, 512, 3, 3) * (1.5 / np.sqrt(512 * 9))).astype(np.float32))
    return sd


def make_eval_dataset(root: str, sal: str, seed: int = 0, method: str = "UAVSal", halve_second: bool = True):
    """A tiny evaluation tree in the layout evalscores_vid_torch walks (utils_score_torch.py:473-490): two 5-frame videos at
    36x64 (12 fixations per frame), the second one's saliency maps at half size (the driver's cv2.resize path).  Written with
    the package's MAT v7.3 writer (halve_second=False keeps both at full size: the `_sum` protocol asserts equal sizes).  Returns
    the arrays for reference."""
    import os
    from iip_uavsal_saliency_b200 import mat73
    H, W, F = 36, 64, 5
    for d in (root + "maps/", root + "fixations/maps/", sal + "Saliency/" + method + "/"):
        os.makedirs(d, exist_ok=True)
    out = {}
    for v, name in enumerate(("vidA", "vidB")):
        pred, true = make_metric_pairs(F, H, W, seed=50 + seed + v)
        fixmap = np.rint(true[:, 0]).astype(np.uint8).transpose(1, 2, 0)[:, :, None, :]           # (H,W,1,F)
        fixpts = np.zeros((H, W, 1, F), np.uint8)
        rs = np.random.RandomState(70 + seed + v)
        for f in range(F):
            p = true[f, 0].ravel().astype(np.float64) + 1e-3
            idx = rs.choice(H * W, size=12, replace=False, p=p / p.sum())
            fixpts[:, :, 0, f].flat[idx] = 1
        sal_u8 = np.rint(pred[:, 0]).astype(np.uint8)
        if v == 1 and halve_second:
            sal_u8 = sal_u8[:, ::2, ::2]
        salmap = np.ascontiguousarray(sal_u8.transpose(1, 2, 0)[:, :, None, :])
        mat73.savemat(sal + "Saliency/" + method + "/" + name + ".mat", {"salmap": salmap})
        mat73.savemat(root + "maps/" + name + "_fixMaps.mat", {"fixMap": fixmap})
        mat73.savemat(root + "fixations/maps/" + name + "_fixPts.mat", {"fixLoc": fixpts})
        out[name] = (salmap, fixmap, fixpts)
    return out
