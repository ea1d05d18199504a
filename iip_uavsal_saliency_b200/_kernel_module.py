"""Shared machinery of the drop-in nn.Modules: a per-shape plan cache keyed on the parameter versions, and
the NCHW-fp32 <-> arena staging used when a sub-module is called on its own."""
from __future__ import annotations

import os
from typing import Callable, Dict, Tuple

import torch
import torch.nn as nn

from .engine import Buf, Plan


def default_terms() -> int:
    return 1 if os.environ.get("UAVSAL_PRECISION", "exact").lower() == "fast" else 3


def default_engine() -> str:
    return os.environ.get("UAVSAL_ENGINE", "tc").lower()


def require_cuda(t: torch.Tensor, who: str):
    if not t.is_cuda:
        raise RuntimeError("%s: uavsal-b200 runs on CUDA (sm_100a) only; got a %s tensor and there is no CPU "
                           "fallback" % (who, t.device))


class KernelModule(nn.Module):
    """nn.Module whose forward is a cached kernel plan.  Sub-classes implement
    ``_emit(plan, x: Buf, n, h, w) -> (Buf, h, w)``."""

    def _plan_cache(self) -> Dict:
        c = self.__dict__.get("_plans")
        if c is None:
            c = {}
            self.__dict__["_plans"] = c
        return c

    def _weights_signature(self) -> Tuple:
        """One (data_ptr, _version) pair per parameter / buffer: any re-allocation (``.to()``, ``.cuda()``) or tracked in-place
        update (``load_state_dict``, ``copy_``) of any tensor gives a new signature, hence new plans and re-packed weights.
        Edits the version counter does not see (``p.data.copy_``, raw-pointer writes) need ``invalidate()``."""
        return tuple((t.data_ptr(), t._version) for t in list(self.parameters()) + list(self.buffers()))

    def invalidate(self):
        """Forget every cached plan and packed weight of this module (and its sub-modules)."""
        from .engine import invalidate_packed
        invalidate_packed(self)
        for m in self.modules():
            m.__dict__.pop("_plans", None)
            h = m.__dict__.get("_holder")
            if h is not None:
                h.__dict__.pop("_plans", None)
        return self

    def _mode(self) -> Tuple[int, str]:
        return (self.__dict__.get("_terms") or default_terms(), self.__dict__.get("_engine") or default_engine())

    def set_mode(self, precision: str = None, engine: str = None):
        """precision: 'exact' (bf16x3 split MMA, default) | 'fast' (bf16x1); engine: 'tc' (tcgen05) | 'simt'."""
        if precision is not None:
            self.__dict__["_terms"] = {"exact": 3, "fast": 1}[precision]
        if engine is not None:
            assert engine in ("tc", "tc1", "simt")
            self.__dict__["_engine"] = engine
        self._plan_cache().clear()
        return self

    def _cached_plan(self, key, builder: Callable[[Plan], None]) -> Plan:
        terms, engine = self._mode()
        full = (key, terms, engine, self._weights_signature())
        cache = self._plan_cache()
        plan = cache.get(full)
        if plan is None:
            # evict by bytes, least recently used first (every distinct call shape / slot owns an arena)
            budget = int(os.environ.get("UAVSAL_PLAN_CACHE_GB", "48")) << 30
            while cache and (len(cache) > 64 or sum(p.arena_bytes for p in cache.values()) > budget):
                cache.pop(next(iter(cache)))
            plan = Plan.build(key[0], terms, engine, builder)
            if plan.device.type == "cuda":
                torch.cuda.current_stream(plan.device).synchronize()       # weight packing is done before any stream replays the plan
        else:
            cache.pop(full)                                                # re-insert: dict order = recency
        cache[full] = plan
        return plan

    # generic single-input, single-output NCHW forward
    def _forward_nchw(self, x: torch.Tensor) -> torch.Tensor:
        require_cuda(x, type(self).__name__)
        n, c, h, w = x.shape

        def build(plan: Plan):
            xin = plan.tensor((n, c, h, w))
            xb = plan.alloc(n * h * w, c)
            plan.pack_nchw(xin, n, c, h, w, xb)
            yb, ho, wo = self._emit(plan, xb, n, h, w)
            yout = plan.tensor((n, yb.c, ho, wo))
            plan.unpack_nchw(yb, n, yb.c, ho, wo, yout)
            plan.named.update(x_in=xin, y_out=yout)

        plan = self._cached_plan((x.device, "nchw", n, c, h, w), build)
        plan.named["x_in"].copy_(x)
        plan.launch()
        return plan.named["y_out"].clone()
