"""Run ONE eager 20-frame UAVSal call (config #2 shapes) between cudaProfilerStart/Stop, for ncu:

    ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
        --clock-control none --csv --log-file gpurun_out/call20_kernels.csv python tools/profile_call.py
    python tools/summarize_ncu.py gpurun_out/call20_kernels.csv profiles/r01_call20

Dev tool (synthetic inputs / weights from oracle.synth; nothing is checked here)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

from iip_uavsal_saliency_b200.model import UAVSal
from oracle import synth


def main():
    dev = torch.device("cuda", 0)
    precision = sys.argv[1] if len(sys.argv) > 1 else "exact"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20            # 60 = the whole-clip plan (three reference calls, group 20)
    m = UAVSal().eval()
    m.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
    m = m.to(dev).set_mode(precision=precision)
    clips = max(1, n // 60)                                       # 120 = two clips batched into one plan
    plan = m.get_plan(dev, n, 360, 640, x_kind=2, post_hw=(360, 640), cb_shared=True, group=20 if n > 20 else 0, clips=clips)
    pr = np.load(os.path.join(ROOT, "tests", "golden", "priors.npz"))
    plan.named["cb_gauss_in"].copy_(torch.from_numpy(pr["gauss"].transpose(2, 0, 1)[None]))
    plan.named["cb_ob_in"].copy_(torch.from_numpy((pr["uav2_u8"].astype(np.float32) / 255).transpose(2, 0, 1)[None]))
    for ci in range(clips):
        per = n // clips
        plan.named["x_in"][ci * per:(ci + 1) * per].copy_(torch.from_numpy(synth.make_clip(2 + ci, per, 360, 640)))
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    plan.run()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("ops", len(plan.ops), "launches", plan.num_launches, "tags", ",".join(op.tag for op in plan.ops))


if __name__ == "__main__":
    main()
