"""Are pair-mode (cta_group::2) and single-CTA tcgen05 GEMM results bit-identical?  And fused vs unfused block variants? (dev tool)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iip_uavsal_saliency_b200 import _ext
from iip_uavsal_saliency_b200.engine import Plan

dev = torch.device("cuda")
lib = _ext.load()
torch.manual_seed(0)
outs = {}
for cl in (1, 2):
    lib.uavsal_set_option(4, cl)
    for (m, k, n) in [(72000, 256, 1536), (72000, 1536, 256)]:
        torch.manual_seed(1)
        p = Plan(dev, 3, "tc")
        a = p.alloc(m, k); a.t.normal_()
        o = p.alloc(m, n)
        w = torch.randn(n, k, device=dev) / k ** 0.5
        p.pw(a, m, w, torch.zeros(n, device=dev), 1, o)
        p.run(); torch.cuda.synchronize()
        outs[(cl, m, k, n)] = o.t.clone()
for key in [(72000, 256, 1536), (72000, 1536, 256)]:
    a, b = outs[(1,) + key], outs[(2,) + key]
    print(key, "pair == single bitwise:", torch.equal(a, b), "max diff", (a.float() - b.float()).abs().max().item())
