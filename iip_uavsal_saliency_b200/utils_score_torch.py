"""Drop-in for the four saliency metrics of the reference's utils_score_torch.py (metric_cc / metric_nss /
metric_kl / metric_sim, :180-218, with the reduction helpers :20-50 and EPS :13).

All four metrics of a batch come out of ONE fused sm_100a kernel launch (csrc/metrics.cu); the per-metric
functions keep the reference signatures ``metric_x(y_pred (N,1,H,W), y_true (N,2,H,W)) -> (N,1)`` and share the
launch through a one-entry memo, because the reference's evaluation loop calls them back to back on the same
tensors (utils_score_torch.py:551-561).  AUC-Judd / Borji / shuffled (csrc/auc.cu) follow below; the file-walking
evaluation loop stays with the caller.
"""
from __future__ import annotations

import ctypes

import torch

from . import _ext

EPS = 2.2204e-16
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
keys_order = ["AUC_shuffled", "NSS", "AUC_Judd", "AUC_Borji", "KLD", "SIM", "CC"]

_memo = {"key": None, "val": None}


def metrics4(y_pred: torch.Tensor, y_true: torch.Tensor) -> torch.Tensor:
    """(N,4) fp32 columns CC, NSS, KLD, SIM.  Inputs: fp32 or uint8 CUDA tensors, (N,1,H,W) and (N,2,H,W)."""
    if not (y_pred.is_cuda and y_true.is_cuda):
        raise RuntimeError("uavsal-b200 metrics run on CUDA (sm_100a) only; there is no CPU fallback")
    if y_pred.dim() != 4 or y_true.dim() != 4 or y_pred.shape[1] != 1 or y_true.shape[1] != 2 or \
            y_pred.shape[0] != y_true.shape[0] or y_pred.shape[2:] != y_true.shape[2:]:
        raise ValueError("expected y_pred (N,1,H,W) and y_true (N,2,H,W), got %s and %s" % (tuple(y_pred.shape), tuple(y_true.shape)))
    if y_pred.dtype != y_true.dtype or y_pred.dtype not in (torch.float32, torch.uint8):
        y_pred, y_true = y_pred.float(), y_true.float()
    y_pred, y_true = y_pred.contiguous(), y_true.contiguous()
    key = (y_pred.data_ptr(), y_pred._version, y_true.data_ptr(), y_true._version, tuple(y_pred.shape), y_pred.dtype)
    if _memo["key"] == key:
        return _memo["val"]
    n, _, h, w = y_pred.shape
    out = torch.empty((n, 4), dtype=torch.float32, device=y_pred.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(y_pred.device).cuda_stream)
    step = 32768
    for i in range(0, n, step):
        m = min(step, n - i)
        _ext.call("uavsal_metrics4", y_pred[i:i + m].data_ptr(), y_true[i:i + m].data_ptr(),
                  0 if y_pred.dtype == torch.float32 else 1, m, h, w, None, out[i:i + m].data_ptr(), stream)
    _memo["key"], _memo["val"] = key, out
    return out


def metric_cc(y_pred, y_true):
    return metrics4(y_pred, y_true)[:, 0:1]


def metric_nss(y_pred, y_true):
    return metrics4(y_pred, y_true)[:, 1:2]


def metric_kl(y_pred, y_true):
    return metrics4(y_pred, y_true)[:, 2:3]


def metric_sim(y_pred, y_true):
    return metrics4(y_pred, y_true)[:, 3:4]


# ---------------------------------------------------------------------------------------------------
# AUC metrics (utils_score_torch.py:53-177).  The kernels (csrc/auc.cu) do the counting; the random draws stay on the host
# and consume the SAME global generators in the SAME order as the reference (torch.rand for the jitter :82, np.random.randint
# for the sampled pixels :103 / :143), so a seeded run reproduces the reference's scores.
# ---------------------------------------------------------------------------------------------------
def _auc_inputs(y_pred, y_true):
    if y_pred.dim() != 4 or y_true.dim() != 4 or y_pred.shape[1] != 1 or y_true.shape[1] != 2 or \
            y_pred.shape[0] != y_true.shape[0] or y_pred.shape[2:] != y_true.shape[2:]:
        raise ValueError("expected y_pred (N,1,H,W) and y_true (N,2,H,W), got %s and %s" % (tuple(y_pred.shape), tuple(y_true.shape)))
    dev = y_pred.device if y_pred.is_cuda else (y_true.device if y_true.is_cuda else torch.device("cuda"))
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise RuntimeError("uavsal-b200 metrics run on CUDA (sm_100a) only; there is no CPU fallback")
    return y_pred.to(dev, torch.float32).contiguous(), y_true.to(dev, torch.float32).contiguous(), dev


def metric_auc_j(y_pred, y_true, jitter=1):
    """utils_score_torch.py:77-88 -> (N,1); NaN for an empty map / fixation set (:54)."""
    if jitter == True:  # noqa: E712  (the reference's comparison, :81)
        y_pred = y_pred + (torch.rand(y_pred.shape) * 1e-7).to(y_pred.device)
    p, t, dev = _auc_inputs(y_pred, y_true)
    n, _, h, w = p.shape
    out = torch.empty((n,), dtype=torch.float32, device=dev)
    _ext.call("uavsal_auc_judd", p.data_ptr(), t.data_ptr(), n, h, w, out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return out.unsqueeze(1)


def _auc_sampled(p, t, dev, draw, n_rep=100, step=0.1):
    """draw(i, n_fix) -> int array (k, n_rep) of flat pixel indices for pair i (or None: the reference returns NaN before drawing)."""
    import numpy as np
    n, _, h, w = p.shape
    flat = p.flatten(1)
    valid = ((flat.amax(1) > flat.amin(1)) & torch.isfinite(flat.amax(1))).cpu().numpy()       # any(S > 0), :92
    n_fix = (t[:, 1].flatten(1) > 0.5).sum(1).cpu().numpy()
    draws = [draw(i, int(n_fix[i])) if (valid[i] and n_fix[i] > 0) else None for i in range(n)]
    max_k = max([d.shape[0] for d in draws if d is not None] + [1])
    idx = np.zeros((n, max_k, n_rep), np.int32)
    n_k = np.zeros((n,), np.int32)
    for i, d in enumerate(draws):
        if d is not None:
            idx[i, :d.shape[0]] = d
            n_k[i] = d.shape[0]
    idx_d, nk_d = torch.from_numpy(idx).to(dev), torch.from_numpy(n_k).to(dev)
    out = torch.empty((n,), dtype=torch.float32, device=dev)
    _ext.call("uavsal_auc_sampled", p.data_ptr(), t.data_ptr(), n, h, w, idx_d.data_ptr(), nk_d.data_ptr(), max_k, n_rep, float(step),
              out.data_ptr(), ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    return out.unsqueeze(1)


def metric_auc_b(y_pred, y_true):
    """utils_score_torch.py:123-132 (auc_b :91-120) -> (N,1)."""
    import numpy as np
    p, t, dev = _auc_inputs(y_pred, y_true)
    n_pixels = p.shape[2] * p.shape[3]
    return _auc_sampled(p, t, dev, lambda i, n_fix: np.random.randint(0, n_pixels, [n_fix, 100]))


def metric_auc_s(y_pred, y_true, shuff_map):
    """utils_score_torch.py:162-172 (auc_s :135-159) -> (N,1).  shuff_map (N,1,H,W): other frames' fixation counts."""
    import numpy as np
    p, t, dev = _auc_inputs(y_pred, y_true)
    oth = torch.flatten(shuff_map, 1, -1).cpu().numpy()

    def draw(i, n_fix):
        ind = np.nonzero(oth[i])[0]
        n_ind = len(ind)
        r = np.random.randint(0, n_ind, [n_ind, 100])[:min(n_fix, n_ind), :]       # the reference draws n_ind rows and keeps n_fix_oth (:143)
        return ind[r]

    return _auc_sampled(p, t, dev, draw)


metrics = {"AUC_shuffled": metric_auc_s, "AUC_Judd": metric_auc_j, "AUC_Borji": metric_auc_b,
           "NSS": metric_nss, "CC": metric_cc, "SIM": metric_sim, "KLD": metric_kl}


# reduction helpers of the reference (utils_score_torch.py:20-50): per-(n,c) scalars broadcast back to (H,W).
# The fused kernel never materialises these; they are kept for API compatibility.
def _bcast(v, like):
    return v.expand(-1, -1, like.shape[2], like.shape[3]).contiguous()


def get_sum(input):
    return _bcast(torch.sum(input, (2, 3), keepdim=True), input)


def get_max(input):
    return _bcast(torch.amax(input, (2, 3), keepdim=True), input)


def get_min(input):
    return _bcast(torch.amin(input, (2, 3), keepdim=True), input)


def get_mean(input):
    return _bcast(torch.mean(input, (2, 3), keepdim=True), input)


def get_std(input):
    return _bcast(torch.std(input, (2, 3), keepdim=True), input)
