"""Drop-in for the reference's entry script: ``Demo_Test.test`` (/root/reference/Demo_Test.py:30-95) and ``get_bias`` (:14-27).

Same arguments, same outputs - one ``<video>.mat`` per input video holding ``salmap`` uint8 (H, W, 1, F) at the video's
native size - with the whole per-video loop on the device:

    cv2 decode (host)  ->  uavsal_letterbox_u8  ->  ClipRunner (stem .. post-process)  ->  mat73.savemat

Differences a user sees: ``model_path`` may be any file ``checkpoint.load_reference_state_dict`` reads (the reference's
whole-module pickle, or a state dict); the prior ``.mat`` files are looked up in ``priors_path`` (default: the current
directory, as the reference does); ``DataSet_Train`` is an argument instead of a module global.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from . import mat73
from .model import UAVSal
from .runner import ClipRunner
from .utils_data import get_guasspriors, get_ob_priors, preprocess_videos

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


def get_bias(bias_type=[1, 1, 1], batch_size=2, shape_r=45, shape_c=80, DataSet_Train="UAV2", priors_path=""):
    """Demo_Test.py:14-27: [gauss (N,8,h,w), ob (N,20,h,w)] prior tensors (empty tensors for disabled branches)."""
    if bias_type[0]:
        g = torch.tensor(get_guasspriors(batch_size, shape_r, shape_c, channels=8, priors_path=priors_path).transpose((0, 3, 1, 2))).float()
    else:
        g = torch.tensor([]).float()
    if bias_type[1]:
        o = torch.tensor(get_ob_priors("", DataSet_Train, "train", batch_size, shape_r, shape_c, priors_path=priors_path).transpose((0, 3, 1, 2))).float()
    else:
        o = torch.tensor([]).float()
    return [g.to(device), o.to(device)]


def test(input_path, output_path, model_path, method_name="UAVSal", saveFrames=float("inf"), time_dims=5, iosize=[480, 640, 60, 80],
         batch_size=4, bias_type=[1, 1, 1], DataSet_Train="UAV2", priors_path=""):
    """Demo_Test.py:30-95.  Returns the list of files written."""
    if device.type != "cuda":
        raise RuntimeError("uavsal-b200 inference runs on CUDA (sm_100a) only; there is no CPU fallback")
    if list(bias_type) != [1, 1, 1]:
        raise NotImplementedError("the device runner takes both prior maps (bias_type=[1,1,1], the published configuration); "
                                  "other combinations run through UAVSal.forward")
    model = UAVSal(cnn_type="mobilenet_v2", time_dims=time_dims, num_stblock=2, bias_type=bias_type, iosize=iosize, planes=256,
                   pre_model_path="")
    if not os.path.exists(model_path):
        raise ValueError                                            # Demo_Test.py:40-41
    model.load_reference(model_path)
    model = model.to(device).eval()
    output_path = output_path + method_name + "/"
    os.makedirs(output_path, exist_ok=True)
    shape_r, shape_c, shape_r_out, shape_c_out = iosize
    gauss = get_guasspriors(1, shape_r_out, shape_c_out, channels=8, priors_path=priors_path)[0].astype(np.float32)
    ob = get_ob_priors("", DataSet_Train, "train", 1, shape_r_out, shape_c_out, priors_path=priors_path)[0].astype(np.float32)
    runners = {}
    written = []
    names = sorted(f for f in os.listdir(input_path) if f.endswith(".avi") or f.endswith(".AVI") or f.endswith(".mp4"))
    for name in names:
        ovideo_path = output_path + name[:-4] + ".mat"
        if os.path.exists(ovideo_path):
            continue
        frames, nframes, height, width = preprocess_videos(input_path + name, shape_r, shape_c, saveFrames, mode="RGB", normalize=False,
                                                           device=device)
        keep = (nframes // time_dims) * time_dims                   # frames beyond the last full chunk are never produced (:68-70)
        if keep == 0:
            mat73.savemat(ovideo_path, {"salmap": np.zeros((height, width, 1, 0), np.uint8)})
            written.append(ovideo_path)
            continue
        key = (height, width)
        if key not in runners:
            runners[key] = ClipRunner(model, gauss, ob, batch_size=batch_size, out_hw=key)
        _, u8 = runners[key].run_clip(frames, want_maps=False)
        pred = u8[:int(min(keep, saveFrames))].cpu().numpy()        # (F, H, W)
        mat73.savemat(ovideo_path, {"salmap": np.ascontiguousarray(pred.transpose(1, 2, 0)[:, :, None, :])})
        written.append(ovideo_path)
    return written
