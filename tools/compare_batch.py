"""Per-op CUDA-event times of the SRF-Net plan at 20 vs 60 frames (dev tool)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iip_uavsal_saliency_b200.model import UAVSal
from oracle import synth

dev = torch.device("cuda", 0)
m = UAVSal().eval()
m.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
m = m.to(dev)
res = {}
for n in (20, 60):
    plan = m.get_plan(dev, n, 360, 640, x_kind=2, stage="sfnet")
    plan.named["x_in"].copy_(torch.from_numpy(synth.make_clip(2, n, 360, 640)))
    plan.run(); torch.cuda.synchronize()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for rep in range(2):
        evs = []
        for op in plan.ops:
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record(); op.fn(*op.args, stream); e1.record()
            evs.append((op.tag, e0, e1))
        torch.cuda.synchronize()
    res[n] = [(t, a.elapsed_time(b) * 1e3) for t, a, b in evs]
tot20 = tot60 = 0
for (t, a), (_, b) in zip(res[20], res[60]):
    tot20 += a; tot60 += b
    print("%-24s %8.1f %8.1f  x%.2f" % (t, a, b, b / a))
print("TOTAL %.1f %.1f x%.2f" % (tot20, tot60, tot60 / tot20))
