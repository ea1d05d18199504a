"""Drop-in for the four saliency metrics of the reference's utils_score_torch.py (metric_cc / metric_nss /
metric_kl / metric_sim, :180-218, with the reduction helpers :20-50 and EPS :13).

All four metrics of a batch come out of ONE fused sm_100a kernel launch (csrc/metrics.cu): ``metrics4`` returns the (N,4)
table.  The per-metric functions keep the reference signatures ``metric_x(y_pred (N,1,H,W), y_true (N,2,H,W)) -> (N,1)``;
each is its own launch of the fused kernel (no result is cached between calls: tensors of successive batches reuse
addresses, so nothing cheap identifies "the same inputs").  Callers that want all four - the evaluation driver below -
call ``metrics4`` once per batch.  AUC-Judd / Borji / shuffled (csrc/auc.cu) follow below.
"""
from __future__ import annotations

import ctypes

import torch

from . import _ext

EPS = 2.2204e-16
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
keys_order = ["AUC_shuffled", "NSS", "AUC_Judd", "AUC_Borji", "KLD", "SIM", "CC"]


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
    if y_pred.dtype == torch.float32:                 # float4 loads: a view into the middle of a buffer is re-based
        y_pred = y_pred.clone() if y_pred.data_ptr() % 16 else y_pred
        y_true = y_true.clone() if y_true.data_ptr() % 16 else y_true
    n, _, h, w = y_pred.shape
    out = torch.empty((n, 4), dtype=torch.float32, device=y_pred.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(y_pred.device).cuda_stream)
    step = 32768
    for i in range(0, n, step):
        m = min(step, n - i)
        _ext.call("uavsal_metrics4", y_pred[i:i + m].data_ptr(), y_true[i:i + m].data_ptr(),
                  0 if y_pred.dtype == torch.float32 else 1, m, h, w, None, out[i:i + m].data_ptr(), stream)
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
    # frames with more than 4096 fixations are sorted in this workspace instead of shared memory (the reference has no cap, :53-74)
    ws_bytes = int(_ext.load().uavsal_auc_judd_workspace(n, h, w))
    ws = torch.empty((max(ws_bytes, 16),), dtype=torch.uint8, device=dev)
    _ext.call("uavsal_auc_judd", p.data_ptr(), t.data_ptr(), n, h, w, out.data_ptr(), ws.data_ptr(), ws_bytes,
              ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    ws.record_stream(torch.cuda.current_stream(dev))
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


# ---------------------------------------------------------------------------------------------------
# evaluation driver (utils_score_torch.py:231-582): directory walking and .mat I/O on the host (mat73 instead of hdf5storage),
# all seven metrics on the device.  Random draws (shuffle maps :334, sampled AUCs) consume the global numpy generator in the
# reference's order, so a seeded run reproduces the reference's score files.
# ---------------------------------------------------------------------------------------------------
shuff_size = {"SALICON": (480, 640), "DIEM": (480, 640), "DIEM20": (480, 640), "CITIUS": (240, 320), "SFU": (288, 352),
              "LEDOV": (1080, 1920), "LEDOV41": (1080, 1920), "UAV2-TE": (720, 1280), "UAV2": (720, 1280), "default": (480, 640),
              "AVS1K-TE": (720, 1280), "AVS1K": (720, 1280)}


def resize_fixation(img, rows=480, cols=640):
    """utils_score_torch.py:248-263."""
    import numpy as np
    out = np.zeros((rows, cols), np.uint8)
    fr, fc = rows / img.shape[0], cols / img.shape[1]
    for coord in np.argwhere(img):
        r, c = int(np.round(coord[0] * fr)), int(np.round(coord[1] * fc))
        out[r - 1 if r == rows else r, c - 1 if c == cols else c] = 1
    return out


def getALLFix_vid(fixsDir, DataSet="DIEM20", maxframes=float("inf")):
    """utils_score_torch.py:300-330: normalised fixation coordinates of every frame of every video."""
    import os
    import numpy as np
    from . import mat73
    names = sorted(f for f in os.listdir(fixsDir) if f.endswith(".mat"))
    num = 45 if DataSet.upper() == "CITIUS" else len(names)
    if DataSet.upper() == "DIEM20":
        maxframes = 300
    pts = []
    for name in names[:num]:
        fix = mat73.loadmat(fixsDir + name)["fixLoc"]
        for f in range(int(min(maxframes, fix.shape[3]))):
            fx, fy = np.where(fix[:, :, 0, f])
            pts.append(np.concatenate((np.expand_dims(fx / fix.shape[0], 1), np.expand_dims(fy / fix.shape[1], 1)), 1))
    return pts


def getshufmap(ALLFixPts, size=(480, 640), nframes=10):
    """utils_score_torch.py:333-357 (np.random.randint on the global generator, :335)."""
    import numpy as np
    nframes = min(nframes, len(ALLFixPts))
    idx = np.random.randint(0, len(ALLFixPts), int(nframes))
    fix = np.concatenate([ALLFixPts[i] for i in idx], 0).astype(np.float64)      # a copy: the reference scales its list entry in place
    fix[:, 0] *= size[0]
    fix[:, 1] *= size[1]
    fix = np.round(fix).astype(np.int64)
    fix = fix[(fix[:, 0] < size[0]) * (fix[:, 1] < size[1])]
    out = np.zeros(size, dtype=np.uint8)
    out[fix[:, 0], fix[:, 1]] = 1
    return out


def _score_frames(salmap, fixmap, nframes, keys_order, batch_size, shuf_of):
    """The metric loop shared by the two evaluation drivers (utils_score_torch.py:541-572 / 436-466): salmap (F,1,H,W), fixmap
    (F,2,H,W) numpy -> (F, len(keys_order)) scores, NaN rows for frames without a prediction or without ground truth.
    ``shuf_of(itrue)`` supplies AUC_shuffled's third argument for a batch."""
    import math
    import numpy as np
    iscores = np.zeros((nframes, len(keys_order)))
    # CC / NSS / KLD / SIM come out of one fused launch per batch; they draw no random numbers, so computing them ahead of
    # the reference's metric-major loop (:541-561) leaves the generator order of the sampled AUCs untouched
    fused = [m for m in ("CC", "NSS", "KLD", "SIM") if m in keys_order]
    if fused:
        col = {"CC": 0, "NSS": 1, "KLD": 2, "SIM": 3}
        for b in range(math.ceil(nframes / batch_size)):
            ipred = torch.tensor(salmap[b * batch_size:(b + 1) * batch_size]).float()
            itrue = torch.tensor(fixmap[b * batch_size:(b + 1) * batch_size]).float()
            m4 = metrics4(ipred.to(device), itrue.to(device)).cpu()
            for m in fused:
                iscores[b * batch_size:(b + 1) * batch_size, keys_order.index(m)] = m4[:, col[m]]
    for k, metric in enumerate(keys_order):
        if metric in fused:
            continue
        func = metrics[metric]
        for b in range(math.ceil(nframes / batch_size)):
            ipred = torch.tensor(salmap[b * batch_size:(b + 1) * batch_size]).float()
            itrue = torch.tensor(fixmap[b * batch_size:(b + 1) * batch_size]).float()
            if metric == "AUC_shuffled":
                m = func(ipred, itrue, shuf_of(itrue))
            elif metric == "AUC_Borji":
                m = func(ipred, itrue)
            else:
                m = func(ipred.to(device), itrue.to(device))
            iscores[b * batch_size:(b + 1) * batch_size, k] = m.data.cpu()[:, 0]
    for f in range(nframes):
        if not np.any(salmap[f, 0]) or not np.any(fixmap[f], axis=(1, 2)).all():
            iscores[f] = np.nan
    return iscores


def evalscores_vid_torch(RootDir, SalDir, DataSet, MethodNames, keys_order=keys_order, batch_size=64):
    """utils_score_torch.py:473-582: per-video score files ``Scores/<method>/Score_<video>.mat`` ({'iscore': (frames, metrics)})
    from ``Saliency/<method>/<video>.mat`` (salmap), ``maps/<video>_fixMaps.mat`` and ``fixations/maps/<video>_fixPts.mat``.
    Returns {method: {video: iscores}} (the reference keeps this dict local)."""
    import math
    import os
    import numpy as np
    from . import mat73
    mapsDir, fixsDir = RootDir + "maps/", RootDir + "fixations/maps/"
    salsDir, scoreDir = SalDir + "Saliency/", SalDir + "Scores/"
    os.makedirs(scoreDir, exist_ok=True)
    all_pts = []
    if "AUC_shuffled" in keys_order:
        pts_path = RootDir + "ALLFixPts_" + DataSet.upper() + ".npy"
        if not os.path.exists(pts_path):
            all_pts = getALLFix_vid(fixsDir, DataSet)
            arr = np.empty(len(all_pts), dtype=object)                 # ragged list: an object array, as old numpy made implicitly
            for i, a in enumerate(all_pts):
                arr[i] = a
            np.save(pts_path, arr)
        else:
            all_pts = list(np.load(pts_path, allow_pickle=True))
    result = {}
    for method in MethodNames:
        if os.path.exists(scoreDir + "Score_" + method + ".mat"):
            continue
        iscoreDir = scoreDir + method + "/"
        os.makedirs(iscoreDir, exist_ok=True)
        salmap_dir = salsDir + method + "/"
        scores = {}
        for sal_name in sorted(f for f in os.listdir(salmap_dir) if f.endswith(".mat")):
            file_name = sal_name[:-4]
            iscore_path = iscoreDir + "Score_" + file_name + ".mat"
            if os.path.exists(iscore_path):
                scores[file_name] = mat73.loadmat(iscore_path)["iscore"]
                continue
            salmap = mat73.loadmat(salmap_dir + file_name + ".mat")["salmap"]
            fixmap = mat73.loadmat(mapsDir + file_name + "_fixMaps.mat")["fixMap"]
            fixpts = mat73.loadmat(fixsDir + file_name + "_fixPts.mat")["fixLoc"]
            nframes = min(salmap.shape[3], min(fixpts.shape[3], fixmap.shape[3]))
            if salmap.shape[:2] != fixmap.shape[:2]:
                import cv2
                rs = np.zeros((nframes, 1, fixmap.shape[0], fixmap.shape[1]))
                for i in range(nframes):
                    rs[i, 0] = cv2.resize(salmap[:, :, 0, i], (fixmap.shape[1], fixmap.shape[0]))
                salmap = rs
            else:
                salmap = salmap[:, :, :, :nframes].transpose((3, 2, 0, 1))
            fixmap = np.concatenate((fixmap[:, :, :, :nframes], fixpts[:, :, :, :nframes]), axis=2).transpose((3, 2, 0, 1))
            shuf_of = lambda itrue: torch.tensor(np.array([getshufmap(all_pts, size=fixmap.shape[2:]) for _ in range(itrue.shape[0])])).float().unsqueeze(1)
            iscores = _score_frames(salmap, fixmap, nframes, keys_order, batch_size, shuf_of)
            scores[file_name] = iscores
            mat73.savemat(iscore_path, {"iscore": iscores})
        result[method] = scores
    return result


def getSumFix_vid(fixsDir, DataSet="DIEM20", size=None, maxframes=float("inf")):
    """utils_score_torch.py:266-297: the dataset's summed fixation map (the fixed shuffle map of the `_sum` protocol)."""
    import os
    import numpy as np
    from . import mat73
    DataSet = DataSet.upper()
    if size is None:
        size = shuff_size[DataSet] if DataSet in shuff_size else shuff_size["default"]
    if DataSet == "DIEM20":
        maxframes = 300
    shuf = np.zeros(size)
    for name in sorted(f for f in os.listdir(fixsDir) if f.endswith(".mat")):
        fix = mat73.loadmat(fixsDir + name)["fixLoc"]
        use = int(min(maxframes, fix.shape[3]))
        fix = fix[:, :, :, :use]
        if fix.shape[:2] != tuple(size):
            fix = np.expand_dims(np.array([resize_fixation(fix[:, :, 0, i], size[0], size[1]) for i in range(use)]).transpose((1, 2, 0)), axis=2)
        shuf += np.sum(fix[:, :, 0, :], axis=2)
        shuf = np.round(shuf)
    return shuf


def evalscores_vid_torch_sum(RootDir, SalDir, DataSet, MethodNames, keys_order=keys_order, batch_size=64):
    """utils_score_torch.py:368-470 (the "STRNN" protocol): as evalscores_vid_torch, but AUC-shuffled uses ONE fixed shuffle map -
    the dataset's summed fixation map (``getSumFix_vid``, cached as ``RootDir/Shuffle_<DATASET>.mat``, resized with
    ``resize_fixation`` when its size differs from the fixation maps') - the saliency maps must already have the ground truth's
    size, and the score files go to ``Scores_sum/``.

    Two behaviours of the reference are kept as they are: (1) its ``shuffle_map != []`` test is meant as "a shuffle map is in use"
    (numpy < 1.25 answered True for an array; numpy 2 raises, so the reference cannot run this protocol today); (2) the (H, W) map
    goes to ``metric_auc_s`` unbatched, whose ``flatten(1, -1)`` leaves it (H, W), so frame i of a batch draws its "other" pixels
    from ROW i of the map (:162-172) - hence batch_size must not exceed H.  Pinned against the unmodified reference run with an
    ndarray subclass that restores (1) (tests/golden/eval_driver_sum.npz)."""
    import os
    import numpy as np
    from . import mat73
    mapsDir, fixsDir = RootDir + "maps/", RootDir + "fixations/maps/"
    salsDir, scoreDir = SalDir + "Saliency/", SalDir + "Scores_sum/"
    os.makedirs(scoreDir, exist_ok=True)
    shuffle_map = None
    if "AUC_shuffled" in keys_order:
        shuff_path = RootDir + "Shuffle_" + DataSet.upper() + ".mat"
        if not os.path.exists(shuff_path):
            shuffle_map = getSumFix_vid(fixsDir, DataSet)
            mat73.savemat(shuff_path, {"ShufMap": shuffle_map})
        else:
            shuffle_map = mat73.loadmat(shuff_path)["ShufMap"]
    result = {}
    for method in MethodNames:
        if os.path.exists(scoreDir + "Score_" + method + ".mat"):
            continue
        iscoreDir = scoreDir + method + "/"
        os.makedirs(iscoreDir, exist_ok=True)
        salmap_dir = salsDir + method + "/"
        scores = {}
        for sal_name in sorted(f for f in os.listdir(salmap_dir) if f.endswith(".mat")):
            file_name = sal_name[:-4]
            iscore_path = iscoreDir + "Score_" + file_name + ".mat"
            if os.path.exists(iscore_path):
                scores[file_name] = mat73.loadmat(iscore_path)["iscore"]
                continue
            salmap = mat73.loadmat(salmap_dir + file_name + ".mat")["salmap"]
            fixmap = mat73.loadmat(mapsDir + file_name + "_fixMaps.mat")["fixMap"]
            fixpts = mat73.loadmat(fixsDir + file_name + "_fixPts.mat")["fixLoc"]
            ishuf = shuffle_map
            if shuffle_map is not None and shuffle_map.shape != fixpts.shape[:2]:
                ishuf = resize_fixation(shuffle_map, fixpts.shape[0], fixpts.shape[1])
            ishuf = torch.tensor(np.asarray(ishuf if ishuf is not None else [])).float()
            nframes = min(salmap.shape[3], min(fixpts.shape[3], fixmap.shape[3]))
            salmap = salmap[:, :, :, :nframes].transpose((3, 2, 0, 1))
            fixmap = np.concatenate((fixmap[:, :, :, :nframes], fixpts[:, :, :, :nframes]), axis=2).transpose((3, 2, 0, 1))
            if salmap.shape[2:] != fixmap.shape[2:]:
                raise AssertionError("saliency maps %s and fixation maps %s differ in size (utils_score_torch.py:432)" % (salmap.shape[2:], fixmap.shape[2:]))
            iscores = _score_frames(salmap, fixmap, nframes, keys_order, batch_size, lambda itrue: ishuf)
            scores[file_name] = iscores
            mat73.savemat(iscore_path, {"iscore": iscores})
        result[method] = scores
    return result
