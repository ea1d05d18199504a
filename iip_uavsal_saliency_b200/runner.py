"""Per-clip inference loop with the semantics of Demo_Test.test (Demo_Test.py:68-91), entirely on device:
uint8 frames in -> uint8 saliency maps out.

  * frames beyond floor(F/time_dims)*time_dims are dropped (quirk Q1);
  * a clip is processed in calls of batch_size*time_dims frames (the temporal differences and the context
    prior see exactly that grouping, quirks Q2/Q3); a shorter last call gets its own plan;
  * the ConvTWA hidden state is handed from call to call (Demo_Test.py:75,85-86);
  * priors are one (h,w,C) map broadcast to every frame, as get_bias builds them (Demo_Test.py:14-27).

Whole-clip plans.  Only two operations of the model see the call grouping: the temporal differences (mirrored edges at
the first / last frame of a call, model.py:194-198) and the context prior's repeat interleave (model.py:361).  Both kernels
take the call size as a parameter, and the ConvTWA state simply runs on from one call's last frame to the next call's first,
so with ``whole_clip=True`` (default) ALL kept frames of a clip go through ONE plan (n = 60 for a 64-frame clip, group = 20):
every GEMM / depthwise launch works on three times the rows (better wave quantisation of the persistent kernels, a third of
the launches) and the results are those of the three chained reference calls.

Batching.  With ``whole_clip=False`` the reference's call loop is kept.  The SRF-Net (backbone + ASPP + 3x3 fuse, model.py:139-158) treats every frame independently, so with
``clip_backbone=True`` (default) it runs ONCE over all kept frames of a clip (stage "sfnet" plan) and the per-call plans
(stage "head": ST blocks, prior fusion, ConvTWA, readout) read their 20-frame slices of its output.  The 12x20 / 23x40
layers of MobileNetV2 are launch-latency-bound at 20 frames; at 60 they do three times the work in about the same time.
Per-frame results are unchanged (eval-mode BN, no cross-frame op before the ST blocks).

Scheduling.  Only the recurrent tail of a call (ConvTWA -> readout -> post-process, the plan's "back" part) depends on
the previous call; everything before it (the "front": backbone, SRF-Net, ST blocks, prior fusion) does not.  The runner
therefore keeps ``depth`` plan instances (own arenas) and runs fronts on ``depth`` CUDA streams while one "back" stream
walks the calls in order, so the latency-bound recurrence of call i overlaps the front of call i+1 - also across clip
boundaries when the caller passes ``sync=False``.  Results are bit-identical to the serial order.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import numpy as np
import torch

from .model import UAVSal


class ClipRunner:
    def __init__(self, model: UAVSal, gauss: np.ndarray, ob: np.ndarray, batch_size: int = 4, out_hw: Optional[Tuple[int, int]] = None,
                 use_graph: bool = True, frame_layout: str = "nhwc", depth: int = 2, clip_backbone: bool = True,
                 single_stream: bool = False, whole_clip: bool = True, clips_per_plan: int = 1, max_plan_frames: int = 240):
        """gauss (h,w,8) / ob (h,w,20) float32 prior maps; frame_layout 'nhwc' (decoder layout) or 'nchw';
        depth = calls in flight (1 = strictly serial on the caller's stream order); clip_backbone: run the SRF-Net once per
        clip instead of once per call."""
        self.model = model
        self.dev = next(model.parameters()).device
        if self.dev.type != "cuda":
            raise RuntimeError("ClipRunner needs the model on a CUDA (sm_100a) device; there is no CPU fallback")
        self.T = model.time_dims
        self.per_call = batch_size * self.T
        self.kind = 2 if frame_layout == "nhwc" else 1
        self.out_hw = out_hw
        self.use_graph = use_graph
        self.depth = max(1, int(depth))
        self.gauss = torch.from_numpy(np.ascontiguousarray(gauss.transpose(2, 0, 1)[None])).float().to(self.dev)
        self.ob = torch.from_numpy(np.ascontiguousarray(ob.transpose(2, 0, 1)[None])).float().to(self.dev)
        # single_stream: queue everything on ONE side stream in program order (no overlap between stages)
        one = torch.cuda.Stream(self.dev) if single_stream else None
        self.front_streams = [one or torch.cuda.Stream(self.dev) for _ in range(self.depth)]
        self.back_stream = one or torch.cuda.Stream(self.dev)
        self._slot_free = [None] * self.depth          # event: the slot's previous call has left the back stream
        self._calls = 0
        self.whole_clip = bool(whole_clip)
        # a whole-clip plan's arena grows with the clip (about 36 MB per frame at 360x640): longer clips run as consecutive
        # plans of at most this many frames (a multiple of the call size, so the call grouping is unchanged), the ConvTWA state
        # handed from one to the next exactly as Demo_Test hands it from call to call (Demo_Test.py:85-86)
        self.max_plan_frames = max(self.per_call, (int(max_plan_frames) // self.per_call) * self.per_call)
        # throughput mode: queue `clips_per_plan` equally shaped clips into ONE plan (the ConvTWA then advances them as a batch
        # of sequences - its per-step launches are latency-bound - and every other launch sees that many times the rows).
        # Only run_clip(..., out=buffer, want_maps=False, sync=False) calls are combined; finish() flushes a partial batch.
        self.clips_per_plan = max(1, int(clips_per_plan))
        self._pending = []
        self.clip_backbone = bool(clip_backbone)
        self.bb_stream = one or torch.cuda.Stream(self.dev)
        # results leave on their own stream: on the back stream the device->host copy of clip k (0.5 ms for two clips' maps)
        # would sit in front of clip k+1's recurrence, the one chain nothing else can overlap
        self.out_stream = one or torch.cuda.Stream(self.dev)
        self._sf_free = [[], []]                       # events: the heads of the previous clip on this slot have read its SRF-Net output
        self._clips = 0

    def _plan(self, n, H, W, slot, stage=None, group=0, clips=1):
        post = self.out_hw or (H, W)
        stage = stage or ("head" if self.clip_backbone else "all")
        plan = self.model.get_plan(self.dev, n, H, W, x_kind=self.kind, post_hw=None if stage == "sfnet" else post, cb_shared=True,
                                   slot=slot, stage=stage, group=group, clips=clips)
        if "ready" not in plan.named:
            torch.cuda.synchronize(self.dev)
            if stage != "sfnet":
                if self.model.use_gauss_prior:
                    plan.named["cb_gauss_in"].copy_(self.gauss)
                if self.model.use_ob_prior:
                    plan.named["cb_ob_in"].copy_(self.ob)
                plan.named["h_in"].zero_()
            if self.use_graph:
                plan.capture()
            torch.cuda.synchronize(self.dev)
            plan.named["ready"] = True
        return plan

    def warm(self, n_frames: int, H: int, W: int):
        """Build (and capture) the plans a clip of n_frames will use on every slot, outside any timed region."""
        keep = (n_frames // self.T) * self.T
        if self.whole_clip:
            batched = self.clips_per_plan > 1 and keep % self.per_call == 0
            for slot in range(2):
                if batched:
                    self._plan(keep * self.clips_per_plan, H, W, slot, "all", self.per_call, self.clips_per_plan)
                if not batched or slot == 0:                        # a left-over single clip always runs on slot 0 (arena memory)
                    self._plan(keep, H, W, slot, "all", self.per_call)
            return
        sizes = {min(self.per_call, keep - i * self.per_call) for i in range(math.ceil(keep / self.per_call))}
        for slot in range(self.depth):
            for n in sizes:
                self._plan(n, H, W, slot)
        if self.clip_backbone:
            for sslot in range(2):
                self._plan(keep, H, W, sslot, "sfnet")

    def finish(self):
        """Make the caller's stream wait for everything the runner has queued (needed after run_clip(sync=False))."""
        self._flush()
        cur = torch.cuda.current_stream(self.dev)
        cur.wait_stream(self.back_stream)
        cur.wait_stream(self.bb_stream)
        cur.wait_stream(self.out_stream)
        for s in self.front_streams:
            cur.wait_stream(s)

    def run_clip(self, frames: torch.Tensor, want_maps: bool = True, out: Optional[torch.Tensor] = None, sync: bool = True):
        """frames: uint8 (F,H,W,3) [nhwc] or (F,3,H,W) [nchw], on the device or in (pinned) host memory.
        Returns (maps fp32 (F',1,h,w) or None, u8 (F',H_out,W_out)); ``out`` (device or pinned host uint8 (>=F',H_out,W_out))
        receives the maps instead of a fresh tensor.  With sync=False the call only queues work: call finish() before
        touching the results from the caller's stream."""
        F_ = frames.shape[0]
        H, W = (frames.shape[1], frames.shape[2]) if self.kind == 2 else (frames.shape[2], frames.shape[3])
        keep = (F_ // self.T) * self.T
        if keep == 0:
            # fewer frames than one chunk: Demo_Test's loop body never runs (count_bs = 0, Demo_Test.py:68-76) and the video gets
            # an empty salmap
            mh, mw = self.model._iosize[2], self.model._iosize[3]
            oh, ow = self.out_hw or (H, W)
            m = torch.empty((0, 1, mh, mw), device=self.dev) if want_maps else None
            return m, (out[:0] if out is not None else torch.empty((0, oh, ow), dtype=torch.uint8, device=self.dev))
        if self.whole_clip:
            return self._run_whole(frames, keep, H, W, want_maps, out, sync)
        ncalls = math.ceil(keep / self.per_call)
        plans = []
        for i in range(ncalls):                                     # build before queueing (capture synchronises)
            n = min(keep, (i + 1) * self.per_call) - i * self.per_call
            plans.append(self._plan(n, H, W, (self._calls + i) % self.depth))
        sf = sslot = None
        if self.clip_backbone:
            sslot = self._clips % 2
            self._clips += 1
            sf = self._plan(keep, H, W, sslot, "sfnet")
        cur = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(cur)                                           # inputs / `out` are ordered after the caller's stream
        bs = self.back_stream
        bs.wait_event(ready)
        sf_done = None
        if sf is not None:                                          # SRF-Net over the whole clip, on its own stream
            bbs = self.bb_stream
            bbs.wait_event(ready)
            for ev in self._sf_free[sslot]:
                bbs.wait_event(ev)
            self._sf_free[sslot] = []
            with torch.cuda.stream(bbs):
                sf.named["x_in"].copy_(frames[:keep], non_blocking=True)
                sf.launch()
                sf_done = torch.cuda.Event()
                sf_done.record(bbs)
            mh, mw = sf.named["map_hw"]
        maps, u8s = [], []
        state = None
        done = 0
        for i in range(ncalls):
            slot = self._calls % self.depth
            self._calls += 1
            plan = plans[i]
            nm = plan.named
            chunk = frames[i * self.per_call:min(keep, (i + 1) * self.per_call)]
            n = chunk.shape[0]
            fs = self.front_streams[slot]
            fs.wait_event(ready)
            if self._slot_free[slot] is not None:
                fs.wait_event(self._slot_free[slot])
            with torch.cuda.stream(fs):
                if sf is not None:
                    fs.wait_event(sf_done)
                    r0 = i * self.per_call * mh * mw
                    nm["sf_in"].t.copy_(sf.named["sf_out"].t[:, r0:r0 + n * mh * mw], non_blocking=True)
                    rd = torch.cuda.Event()
                    rd.record(fs)
                    self._sf_free[sslot].append(rd)
                else:
                    nm["x_in"].copy_(chunk, non_blocking=True)
                plan.launch("front")
                fdone = torch.cuda.Event()
                fdone.record(fs)
            bs.wait_event(fdone)
            with torch.cuda.stream(bs):
                if state is None:
                    nm["h_in"].zero_()
                else:
                    nm["h_in"].copy_(state, non_blocking=True)
                plan.launch("back")
                state = nm["h_out"]                                 # read by the next call on this same stream
                if want_maps:
                    maps.append(nm["out"].clone())
                if out is not None:
                    out[done:done + n].copy_(nm["out_u8"], non_blocking=True)
                else:
                    u8s.append(nm["out_u8"].clone())
                free = torch.cuda.Event()
                free.record(bs)
                self._slot_free[slot] = free
            done += n
        with torch.cuda.stream(bs):
            m = torch.cat(maps, 0) if want_maps else None
            u8 = out[:done] if out is not None else torch.cat(u8s, 0)
        if sync:
            cur.wait_stream(bs)
            for t in (m, u8):
                if t is not None and t.is_cuda:
                    t.record_stream(cur)
        return m, u8

    def _flush(self):
        """Launch whatever run_clip has queued for batching (a partial batch runs clip by clip)."""
        pend, self._pending = self._pending, []
        if not pend:
            return
        if len(pend) == self.clips_per_plan:
            self._launch_whole([p[0] for p in pend], pend[0][1], pend[0][2], pend[0][3], False, [p[4] for p in pend])
        else:
            for fr, keep, H, W, out in pend:
                self._launch_whole([fr], keep, H, W, False, [out])

    def _run_whole(self, frames, keep, H, W, want_maps, out, sync):
        """One plan for the whole clip (see the module docstring); front of clip k+1 overlaps the recurrent back of clip k."""
        if keep > self.max_plan_frames:
            return self._run_long(frames, keep, H, W, want_maps, out, sync)
        if self.clips_per_plan > 1 and not sync and not want_maps and out is not None and keep % self.per_call == 0:
            if self._pending and (self._pending[0][1], self._pending[0][2], self._pending[0][3]) != (keep, H, W):
                self._flush()
            self._pending.append((frames, keep, H, W, out))
            if len(self._pending) == self.clips_per_plan:
                self._flush()
            return None, out[:keep]
        self._flush()
        m, u8 = self._launch_whole([frames], keep, H, W, want_maps, [out])
        if sync:
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_stream(self.back_stream)
            cur.wait_stream(self.out_stream)
            for t in (m, u8):
                if t is not None and t.is_cuda:
                    t.record_stream(cur)
        return m, u8

    def _run_long(self, frames, keep, H, W, want_maps, out, sync):
        """A clip longer than ``max_plan_frames``: consecutive whole-clip plans chained through the ConvTWA state."""
        self._flush()
        maps, u8s, state, done = [], [], None, 0
        while done < keep:
            seg = min(self.max_plan_frames, keep - done)
            m, u8 = self._launch_whole([frames[done:done + seg]], seg, H, W, want_maps, [out[done:done + seg] if out is not None else None],
                                       state=state, keep_state=True)
            state = self._last_state
            maps.append(m)
            u8s.append(u8)
            done += seg
        with torch.cuda.stream(self.back_stream):
            m = torch.cat(maps, 0) if want_maps else None
        if out is not None:
            u8 = out[:keep]
        else:
            with torch.cuda.stream(self.back_stream):
                u8 = torch.cat(u8s, 0)
        if sync:
            cur = torch.cuda.current_stream(self.dev)
            cur.wait_stream(self.back_stream)
            cur.wait_stream(self.out_stream)
            for t in (m, u8):
                if t is not None and t.is_cuda:
                    t.record_stream(cur)
        return m, u8

    def _launch_whole(self, clips, keep, H, W, want_maps, outs, state=None, keep_state=False):
        nc = len(clips)
        slot = self._clips % 2 if (nc > 1 or self.clips_per_plan == 1) else 0
        self._clips += 1
        plan = self._plan(keep * nc, H, W, slot, "all", self.per_call, nc)
        nm = plan.named
        cur = torch.cuda.current_stream(self.dev)
        ready = torch.cuda.Event()
        ready.record(cur)
        fs, bs = self.front_streams[slot % self.depth], self.back_stream
        fs.wait_event(ready)
        bs.wait_event(ready)
        for ev in self._sf_free[slot]:
            fs.wait_event(ev)
        with torch.cuda.stream(fs):
            for ci, fr in enumerate(clips):
                nm["x_in"][ci * keep:(ci + 1) * keep].copy_(fr[:keep], non_blocking=True)
            plan.launch("front")
            fdone = torch.cuda.Event()
            fdone.record(fs)
        bs.wait_event(fdone)
        with torch.cuda.stream(bs):
            if state is None:
                nm["h_in"].zero_()                                  # every clip starts from the zero state (Demo_Test.py:75)
            else:
                nm["h_in"].copy_(state, non_blocking=True)          # continuation of a long clip: the previous plan's last h
            plan.launch("back")
            if keep_state:
                self._last_state = nm["h_out"]                      # read by the next segment on this same (back) stream
            m = nm["out"].clone() if want_maps else None
            u8 = None
            for ci, out in enumerate(outs):
                if out is None:
                    u8 = nm["out_u8"][ci * keep:(ci + 1) * keep].clone()
            bdone = torch.cuda.Event()
            bdone.record(bs)
        frees = [bdone]
        if any(out is not None for out in outs):
            os_ = self.out_stream
            os_.wait_event(bdone)
            with torch.cuda.stream(os_):
                for ci, out in enumerate(outs):
                    if out is not None:
                        out[:keep].copy_(nm["out_u8"][ci * keep:(ci + 1) * keep], non_blocking=True)
                        u8 = out[:keep]
                copied = torch.cuda.Event()
                copied.record(os_)
            frees.append(copied)                                    # the slot's out_u8 may be overwritten only after the copy
        self._sf_free[slot] = frees
        return m, u8
