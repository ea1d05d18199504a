"""Drop-in for the four saliency metrics of the reference's utils_score_torch.py (metric_cc / metric_nss /
metric_kl / metric_sim, :180-218, with the reduction helpers :20-50 and EPS :13).

All four metrics of a batch come out of ONE fused sm_100a kernel launch (csrc/metrics.cu); the per-metric
functions keep the reference signatures ``metric_x(y_pred (N,1,H,W), y_true (N,2,H,W)) -> (N,1)`` and share the
launch through a one-entry memo, because the reference's evaluation loop calls them back to back on the same
tensors (utils_score_torch.py:551-561).  AUC-Judd/Borji/shuffled and the file-walking evaluation loops are
out of scope.
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


metrics = {"NSS": metric_nss, "CC": metric_cc, "SIM": metric_sim, "KLD": metric_kl}


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
