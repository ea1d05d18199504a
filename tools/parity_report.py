"""Parity margins of one Demo_Test-sized call (B=4, T=5, 360x640, lively weights) against the reference-generated golden trace
(tests/golden/call20_trace.npz): per-stage relative L2 error on the sampled taps, max-abs error and CC of the saliency map, with
the widest hidden tensors as q16 rows (default) and as fp32 rows (UAVSAL_HIDDEN_Q16=0).  Dev tool; the same checks, with
thresholds, are tests/test_gpu_parity.py::test_uavsal_call_of_20_frames_vs_reference_golden.

    python tools/parity_report.py > profiles/r02_parity_report.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from oracle import cpu_ref, synth
from oracle.make_golden import sample_idx


def run(q16: bool):
    os.environ["UAVSAL_HIDDEN_Q16"] = "1" if q16 else "0"
    from iip_uavsal_saliency_b200.model import UAVSal
    gold = os.path.join(ROOT, "tests", "golden")
    g = np.load(os.path.join(gold, "call20_trace.npz"))
    pr = np.load(os.path.join(gold, "priors.npz"))
    gauss, ob = pr["gauss"], pr["uav2_u8"].astype(np.float32) / 255
    m = UAVSal().eval()
    m.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
    m = m.cuda()
    x = torch.from_numpy(cpu_ref.normalize_data(synth.make_clip(1, 20, 360, 640).transpose(0, 3, 1, 2))).cuda()
    cb = [torch.from_numpy(np.repeat(gauss.transpose(2, 0, 1)[None], 20, 0).copy()).cuda(),
          torch.from_numpy(np.repeat(ob.transpose(2, 0, 1)[None], 20, 0).copy()).cuda()]
    h0 = torch.from_numpy(np.random.RandomState(7).randn(1, 256, 45, 80).astype(np.float32) * 0.5).cuda()
    plan = m.get_plan(x.device, 20, 360, 640, 0, None, True, False)
    nm = plan.named
    nm["x_in"].copy_(x); nm["cb_gauss_in"].copy_(cb[0]); nm["cb_ob_in"].copy_(cb[1]); nm["h_in"].copy_(h0)
    plan.run()
    torch.cuda.synchronize()
    n_q16 = sum(1 for o in plan.ops if o.name == "uavsal_pw_gemm" and o.args[9] & 32)
    print("hidden rows: %s (%d expand GEMMs write q16)" % ("q16 for >= 1152 channels" if q16 else "fp32 everywhere", n_q16))
    for name in ("c3", "c4", "c5", "sfnet", "st_layer.0", "st_layer.1", "fust", "cb_gauss", "cb_ob", "fucb", "fucbst", "rnn"):
        if name not in nm["taps"] or "trace_" + name not in g.files:
            continue
        buf, hh, ww = nm["taps"][name]
        v = buf.to_float().reshape(20, hh, ww, buf.c).permute(0, 3, 1, 2).contiguous().cpu().numpy()
        ref = g["trace_" + name]
        mine = v.ravel()[sample_idx(v.size, name)]
        print("  %-12s rel-L2 %.2e   (test threshold 5e-4)" % (name, np.linalg.norm(mine - ref) / np.linalg.norm(ref)))
    o = nm["out"].cpu().numpy()
    cc = cpu_ref.metric_cc(torch.from_numpy(o), torch.cat([torch.from_numpy(g["out"])] * 2, 1))
    print("  saliency map  max-abs %.2e (tolerance 2e-3)   min CC %.7f (>= 0.999)" % (np.abs(o - g["out"]).max(), cc.min().item()))


if __name__ == "__main__":
    run(True)
    run(False)
