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
        v = 0
        ptr = 0
        for t in list(self.parameters()) + list(self.buffers()):
            v += t._version
            ptr ^= t.data_ptr()
        return (v, ptr)

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
            if len(cache) > 32:
                cache.clear()
            dev = key[0]
            plan = Plan(dev, terms=terms, engine=engine)
            builder(plan)
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
