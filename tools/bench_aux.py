"""Throughput of the widened-path kernels (video front-end, AUC metrics) on one B200: CUDA events, inputs resident in HBM and
larger than L2.  Dev tool;  python tools/bench_aux.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from iip_uavsal_saliency_b200 import utils_data as ud
from iip_uavsal_saliency_b200 import utils_score_torch as us
from oracle import synth


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6551.4) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6551.4
    for sh, sw, n in [(720, 1280, 128), (1080, 1920, 64)]:
        fr = torch.randint(0, 256, (n, sh, sw, 3), dtype=torch.uint8, device="cuda")
        ms = timed(lambda: ud.letterbox_frames(fr, 360, 640))
        by = n * (sh * sw * 3 + 360 * 640 * 3)
        print(json.dumps({"kernel": "uavsal_letterbox_u8", "src": [sh, sw], "dst": [360, 640], "frames": n, "ms": round(ms, 3),
                          "frames_per_s": round(n / ms * 1e3), "GBps": round(by / ms / 1e6, 1), "hbm_frac": round(by / ms / 1e6 / peak, 3)}))
    pred, true, shuf = synth.make_auc_case(0, n=30)
    p = torch.from_numpy(np.tile(pred, (8, 1, 1, 1))).cuda()
    t = torch.from_numpy(np.tile(true, (8, 1, 1, 1))).cuda()
    n = p.shape[0]
    ms = timed(lambda: us.metric_auc_j(p, t, jitter=0))
    by = n * 2 * 360 * 640 * 4 * 1.0
    print(json.dumps({"kernel": "uavsal_auc_judd", "pairs": n, "ms": round(ms, 3), "pairs_per_s": round(n / ms * 1e3),
                      "note": "three passes over the prediction + one over the fixation plane per pair; one CTA per pair",
                      "GBps_min_traffic": round(by / ms / 1e6, 1)}))
    np.random.seed(0)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); us.metric_auc_b(p[:32], t[:32]); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"kernel": "metric_auc_b (host draws + uavsal_auc_sampled)", "pairs": 32, "ms": round(e0.elapsed_time(e1), 3)}))


if __name__ == "__main__":
    main()
