"""Drop-in for the hot-path helpers of the reference's utils_data.py: normalize_data (:43-65), im2uint8 /
np2mat (:68-82), postprocess_predictions (:289-303), st_get_gaussmaps (:391-412), get_guasspriors (:449-469),
read_ob_priors / get_ob_priors (:552-604), padding / preprocess_videos (:255-287, :321-343; decode on the host with cv2 as
the reference does, letterbox resize + BGR->RGB on the device).  Dataset plumbing is out of scope.  The device pipeline (runner.py) fuses normalisation into the stem kernel and the whole post-process
into uavsal_post_u8; the functions here keep the reference's host-facing signatures.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np
import torch

from . import _ext, mat73

EPS = 2.2204e-16
_MEAN = (0.485, 0.456, 0.406)
_STD = (0.229, 0.224, 0.225)


def normalize_data(data, mean=_MEAN, std=_STD):
    """(u8/255 - mean_c)/std_c in float32 for (3,H,W), (N,3,H,W) or (N,T,3,H,W); input is not modified.
    (Layout/dtype conversion at the API boundary; the runner path feeds uint8 frames straight to the stem kernel.)"""
    ims = data.astype(np.float32) / 255.0 if data.dtype == np.uint8 else data.clone()
    axis = {3: 0, 4: 1, 5: 2}.get(len(ims.shape))
    if axis is None:
        raise ValueError
    for c in range(3):
        idx = (slice(None),) * axis + (c,)
        ims[idx] = (ims[idx] - mean[c]) / std[c]
    return ims


def im2uint8(img):
    if img.dtype == np.uint8:
        return img
    return np.rint(np.clip(img, 0, 255)).astype(np.uint8)


def np2mat(img, dtype=np.uint8):
    return im2uint8(img) if dtype == np.uint8 else img.astype(dtype)


def _post(pred_t: torch.Tensor, shape_r: int, shape_c: int, as_u8: bool) -> torch.Tensor:
    if not pred_t.is_cuda:
        raise RuntimeError("postprocess runs on CUDA (sm_100a) only; there is no CPU fallback")
    p = pred_t.contiguous().float()
    if p.dim() == 2:
        p = p[None]
    n, hs, ws = p.shape
    fm = torch.empty((n,), dtype=torch.float32, device=p.device)
    out = torch.empty((n, shape_r, shape_c), dtype=torch.uint8 if as_u8 else torch.float32, device=p.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(p.device).cuda_stream)
    _ext.call("uavsal_post_u8" if as_u8 else "uavsal_post_f32", p.data_ptr(), n, hs, ws, shape_r, shape_c, fm.data_ptr(),
              out.data_ptr(), stream)
    return out


def postprocess_to_uint8(pred: torch.Tensor, shape_r: int, shape_c: int) -> torch.Tensor:
    """Device pipeline: (n,hs,ws) fp32 CUDA maps -> (n,shape_r,shape_c) uint8 == np2mat(postprocess_predictions(.))."""
    return _post(pred, shape_r, shape_c, True)


def postprocess_predictions(pred, shape_r, shape_c):
    """utils_data.py:289-303 signature: one (h,w) map (numpy or CUDA tensor) -> float (shape_r, shape_c) in [0,255]."""
    if isinstance(pred, np.ndarray):
        if not torch.cuda.is_available():
            raise RuntimeError("postprocess_predictions needs a CUDA device (sm_100a); there is no CPU fallback")
        return _post(torch.from_numpy(np.ascontiguousarray(pred)).cuda(), shape_r, shape_c, False)[0].cpu().numpy()
    return _post(pred, shape_r, shape_c, False)[0]


def st_get_gaussmaps(height, width, nb_gaussian):
    """utils_data.py:391-412 (host numpy; evaluated once per run)."""
    e = height / width
    e1 = (1 - e) / 2
    e2 = e1 + e
    sig = e * np.array(np.arange(1, nb_gaussian + 1)) / 16
    x_t = np.repeat(np.linspace(0.0, 1.0, width)[None, :, None], height, 0).repeat(nb_gaussian, 2)
    y_t = np.repeat(np.linspace(e1, e2, height)[:, None, None], width, 1).repeat(nb_gaussian, 2)
    return 1 / (2 * np.pi * sig * sig + EPS) * np.exp(-((x_t - 0.5) ** 2 / (2 * sig ** 2 + EPS) + (y_t - 0.5) ** 2 / (2 * sig ** 2 + EPS)))


def _quirk_q4_resize(ims, shape_r, shape_c):
    """The reference 'resizes' mismatching priors through a uint8 buffer (utils_data.py:460-464, 595-599), which
    truncates [0,1] floats to 0; reproduced as-is."""
    return np.zeros((shape_r, shape_c, ims.shape[2]), np.uint8)


def get_guasspriors(b_s=2, shape_r=45, shape_c=80, channels=8, priors_path=""):
    path = priors_path + "gauss_priors.mat"
    if os.path.exists(path):
        ims = mat73.loadmat(path)["PriorMaps"]
        if ims.shape[0] != shape_r or ims.shape[1] != shape_c:
            ims = _quirk_q4_resize(ims, shape_r, shape_c)
    else:
        ims = st_get_gaussmaps(shape_r, shape_c, channels)
        ims = (ims - np.min(ims, (0, 1))) / (np.max(ims, (0, 1)) - np.min(ims, (0, 1)) + EPS)
        ims = ims.astype(np.float32)
    return np.repeat(np.expand_dims(ims, axis=0), b_s, axis=0)


def read_ob_priors(datapath, DataSet="", phase_gen="train", shape_r=45, shape_c=80, channels=20, priors_path=""):
    if phase_gen == "train":
        path = priors_path + DataSet.upper() + "_ob_priors_train.mat"
    elif phase_gen == "train_val":
        path = priors_path + DataSet.upper() + "_ob_priors_train_val.mat"
    else:
        raise NotImplementedError
    if not os.path.exists(path):
        raise FileNotFoundError("%s not found; building object priors from a dataset is outside the inference path" % path)
    return mat73.loadmat(path)["PriorMaps"]


def get_ob_priors(datapath, DataSet="", phase_gen="train", b_s=2, shape_r=45, shape_c=80, channels=20, priors_path=""):
    ims = read_ob_priors(datapath, DataSet, phase_gen, shape_r, shape_c, priors_path=priors_path)
    if ims.shape[0] != shape_r or ims.shape[1] != shape_c:
        ims = _quirk_q4_resize(ims, shape_r, shape_c)
    return np.repeat(np.expand_dims(ims, axis=0), b_s, axis=0)


# ---------------------------------------------------------------------------------------------------
# video front-end (utils_data.py:255-287, 321-343)
# ---------------------------------------------------------------------------------------------------
def letterbox_frames(frames_bgr, shape_r: int, shape_c: int, mode: str = "RGB") -> torch.Tensor:
    """(n,h,w,3) uint8 frames in cv2's BGR order (numpy, CPU or CUDA tensor) -> (n,shape_r,shape_c,3) uint8 CUDA tensor:
    ``padding(frame, shape_r, shape_c, 3)`` for every frame (+ the RGB reorder), bit-exact with the reference's cv2 path.
    This is the tensor ``ClipRunner.run_clip`` / the stem kernel take directly (normalisation is fused there)."""
    if mode not in ("RGB", "BGR"):
        raise ValueError
    if not torch.cuda.is_available():
        raise RuntimeError("the video front-end runs on CUDA (sm_100a) only; there is no CPU fallback")
    t = torch.from_numpy(np.ascontiguousarray(frames_bgr)) if isinstance(frames_bgr, np.ndarray) else frames_bgr
    if t.dim() != 4 or t.shape[3] != 3 or t.dtype != torch.uint8:
        raise ValueError("expected (n,h,w,3) uint8 frames, got %s %s" % (tuple(t.shape), t.dtype))
    t = t.cuda(non_blocking=True).contiguous()
    n, h, w, _ = t.shape
    out = torch.empty((n, shape_r, shape_c, 3), dtype=torch.uint8, device=t.device)
    _ext.call("uavsal_letterbox_u8", t.data_ptr(), n, h, w, out.data_ptr(), shape_r, shape_c, 1 if mode == "RGB" else 0,
              ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream))
    return out


def padding(img, shape_r=480, shape_c=640, channels=3):
    """utils_data.py:321-343 for 3-channel uint8 images (numpy in, numpy out)."""
    if channels != 3 or img.ndim != 3:
        raise NotImplementedError("only the 3-channel frame path of preprocess_videos is on the device")
    return letterbox_frames(img[None], shape_r, shape_c, mode="BGR")[0].cpu().numpy()


def preprocess_videos(path, shape_r, shape_c, frames=float("inf"), mode="RGB", normalize=True, device=None):
    """utils_data.py:255-287: decode with cv2.VideoCapture (host, as the reference), letterbox + channel order on the device.
    Returns (ims, nframes, height, width) with ims a numpy array like the reference's - or, with ``device='cuda'``, the uint8
    (n,shape_r,shape_c,3) CUDA tensor to hand to the runner (``normalize`` must then be False: the stem kernel normalises)."""
    import cv2
    if mode not in ("RGB", "BGR"):
        raise ValueError
    cap = cv2.VideoCapture(path)
    nframes = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    width, height = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    nframes = int(min(nframes, frames))
    raw = np.zeros((nframes, height, width, 3), np.uint8)
    for i in range(nframes):
        ret, frame = cap.read()
        raw[i] = frame
    cap.release()
    ims = letterbox_frames(raw, shape_r, shape_c, mode)
    if device is not None:
        if normalize:
            raise ValueError("device output is uint8; the normalisation is fused into the stem kernel")
        return ims.to(device), nframes, height, width
    ims = ims.cpu().numpy()
    if normalize:
        mean, std = (_MEAN, _STD) if mode == "RGB" else (_MEAN[::-1], _STD[::-1])
        ims = ims.astype(np.float32) / 255.0
        for c in range(3):
            ims[:, :, :, c] = (ims[:, :, :, c] - mean[c]) / std[c]
    return ims, nframes, height, width
