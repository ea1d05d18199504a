"""CUDA-graph replay times of the staged plans (dev tool): SRF-Net at 20 / 60 frames, head front / back, full call."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from iip_uavsal_saliency_b200.model import UAVSal
from oracle import synth

dev = torch.device("cuda", 0)
m = UAVSal().eval()
m.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
m = m.to(dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


for n in (20, 60, 120):
    p = m.get_plan(dev, n, 360, 640, x_kind=2, stage="sfnet")
    p.named["x_in"].copy_(torch.from_numpy(synth.make_clip(2, n, 360, 640)))
    p.capture()
    t = timeit(p.launch)
    print("sfnet n=%3d: %.3f ms  (%.1f us/frame)" % (n, t, 1e3 * t / n), flush=True)
    del p
    m._plan_cache().clear()
    torch.cuda.empty_cache()
ph = m.get_plan(dev, 20, 360, 640, x_kind=2, post_hw=(360, 640), cb_shared=True, stage="head")
ph.named["sf_in"].t.normal_()
ph.capture()
print("head front: %.3f ms   back: %.3f ms" % (timeit(lambda: ph.launch("front")), timeit(lambda: ph.launch("back"))), flush=True)
pa = m.get_plan(dev, 20, 360, 640, x_kind=2, post_hw=(360, 640), cb_shared=True)
pa.capture()
print("full call : %.3f ms" % timeit(pa.launch), flush=True)
