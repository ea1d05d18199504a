import copy, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from iip_uavsal_saliency_b200 import engine
from iip_uavsal_saliency_b200.blocks import BasicConv2d
torch.manual_seed(2)
c = BasicConv2d(448, 256, 3)
c[1].running_mean.normal_(); c[1].running_var.uniform_(0.3, 3.0); c[1].weight.data.uniform_(0.5, 1.5); c[1].bias.data.normal_()
host = copy.deepcopy(c)
p = engine.Plan("cuda")
for layout in (engine.W_ROWS_F32, engine.W_ROWS_SPLIT):
    wt, b = p.packed(c.cuda().wspec(), layout, 256, 4032)
    rw, rb = host.wspec().pack_reference(layout, 256, 4032)
    wt = wt.cpu()
    neq = (wt != rw)
    print("layout", layout, "mismatches", int(neq.sum()), "of", wt.numel(), "bias eq", torch.equal(b.cpu(), rb))
    if neq.any():
        idx = neq.nonzero()[:5]
        for i in idx:
            i = tuple(i.tolist())
            print(i, float(wt[i]), float(rw[i]))
        if layout == engine.W_ROWS_F32:
            rows = neq.any(1).nonzero().flatten()
            print("rows with mismatches", rows[:10].tolist(), len(rows))
            sc_ref = host[1].weight / torch.sqrt(host[1].running_var + host[1].eps)
            sc_gpu = (c[1].weight / torch.sqrt(c[1].running_var + c[1].eps)).cpu()
            print("scale cpu vs torch-gpu mismatches", int((sc_ref != sc_gpu).sum()))
