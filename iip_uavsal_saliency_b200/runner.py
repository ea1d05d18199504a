"""Per-clip inference loop with the semantics of Demo_Test.test (Demo_Test.py:68-91), entirely on device:
uint8 frames in -> uint8 saliency maps out.

  * frames beyond floor(F/time_dims)*time_dims are dropped (quirk Q1);
  * a clip is processed in calls of batch_size*time_dims frames (the temporal differences and the context
    prior see exactly that grouping, quirks Q2/Q3); a shorter last call gets its own plan;
  * the ConvTWA hidden state is handed from call to call (Demo_Test.py:75,85-86);
  * priors are one (h,w,C) map broadcast to every frame, as get_bias builds them (Demo_Test.py:14-27).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from .model import UAVSal


class ClipRunner:
    def __init__(self, model: UAVSal, gauss: np.ndarray, ob: np.ndarray, batch_size: int = 4, out_hw: Optional[Tuple[int, int]] = None,
                 use_graph: bool = True, frame_layout: str = "nhwc"):
        """gauss (h,w,8) / ob (h,w,20) float32 prior maps; frame_layout 'nhwc' (decoder layout) or 'nchw'."""
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("ClipRunner needs the model on a CUDA (sm_100a) device; there is no CPU fallback")
        self.T = model.time_dims
        self.per_call = batch_size * self.T
        self.kind = 2 if frame_layout == "nhwc" else 1
        self.out_hw = out_hw
        self.use_graph = use_graph
        self.gauss = torch.from_numpy(np.ascontiguousarray(gauss.transpose(2, 0, 1)[None])).float().to(self.dev)
        self.ob = torch.from_numpy(np.ascontiguousarray(ob.transpose(2, 0, 1)[None])).float().to(self.dev)

    def _plan(self, n, H, W):
        post = self.out_hw or (H, W)
        plan = self.model.get_plan(self.dev, n, H, W, x_kind=self.kind, post_hw=post, cb_shared=True)
        if "ready" not in plan.named:
            if self.model.use_gauss_prior:
                plan.named["cb_gauss_in"].copy_(self.gauss)
            if self.model.use_ob_prior:
                plan.named["cb_ob_in"].copy_(self.ob)
            plan.named["h_in"].zero_()
            if self.use_graph:
                plan.capture()
            plan.named["ready"] = True
        return plan

    def run_clip(self, frames: torch.Tensor, want_maps: bool = True):
        """frames: uint8 (F,H,W,3) [nhwc] or (F,3,H,W) [nchw], on the device or in (pinned) host memory.
        Returns (maps fp32 (F',1,h,w) or None, u8 (F',H_out,W_out)) on the device."""
        F_ = frames.shape[0]
        H, W = (frames.shape[1], frames.shape[2]) if self.kind == 2 else (frames.shape[2], frames.shape[3])
        keep = (F_ // self.T) * self.T
        maps, u8s = [], []
        state = None
        for i in range(math.ceil(keep / self.per_call)):
            chunk = frames[i * self.per_call:min(keep, (i + 1) * self.per_call)]
            plan = self._plan(chunk.shape[0], H, W)
            nm = plan.named
            nm["x_in"].copy_(chunk, non_blocking=True)
            if state is None:
                nm["h_in"].zero_()
            else:
                nm["h_in"].copy_(state)
            plan.launch()
            state = nm["h_out"].clone()
            if want_maps:
                maps.append(nm["out"].clone())
            u8s.append(nm["out_u8"].clone())
        return (torch.cat(maps, 0) if want_maps else None), torch.cat(u8s, 0)
