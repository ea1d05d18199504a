"""Stage-by-stage bit comparison: one 120-frame plan (2 clips) vs two 60-frame plans (dev tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from iip_uavsal_saliency_b200.model import UAVSal
from oracle import synth

dev = torch.device("cuda", 0)
m = UAVSal().eval()
m.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
m = m.to(dev)
pr = np.load(os.path.join(ROOT, "tests", "golden", "priors.npz"))
g_in = torch.from_numpy(pr["gauss"].transpose(2, 0, 1)[None]).float().to(dev)
o_in = torch.from_numpy((pr["uav2_u8"].astype(np.float32) / 255).transpose(2, 0, 1)[None]).float().to(dev)
clips = [torch.from_numpy(synth.make_clip(2 + i, 60, 360, 640)).to(dev) for i in range(2)]


def run(plan, frames):
    nm = plan.named
    nm["x_in"].copy_(frames); nm["cb_gauss_in"].copy_(g_in); nm["cb_ob_in"].copy_(o_in); nm["h_in"].zero_()
    plan.run(); torch.cuda.synchronize()
    taps = {k: v[0].to_float().clone() for k, v in nm["taps"].items()}
    taps["out"] = nm["out"].reshape(-1, 1).clone()
    taps["u8"] = nm["out_u8"].reshape(-1, 1).float().clone()
    return taps


big = m.get_plan(dev, 120, 360, 640, x_kind=2, post_hw=(360, 640), taps=True, cb_shared=True, group=20, clips=2)
tb = run(big, torch.cat(clips, 0))
del big
m._plan_cache().clear(); torch.cuda.empty_cache()
small = m.get_plan(dev, 60, 360, 640, x_kind=2, post_hw=(360, 640), taps=True, cb_shared=True, group=20)
parts = [run(small, c) for c in clips]
for k in tb:
    if k in ("cb_gauss", "cb_ob"):
        continue
    cat = torch.cat([p[k] for p in parts], 0)
    d = (tb[k] - cat).abs()
    per_frame = d.reshape(120, -1).amax(1)
    bad = (per_frame > 0).nonzero().flatten().tolist()
    print("%-12s equal=%s maxdiff=%.3e frames differing: %d first %s" % (k, bool((d == 0).all()), d.max().item(), len(bad), bad[:8]), flush=True)
