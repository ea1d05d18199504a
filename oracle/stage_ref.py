"""Stage the reference's own Python sources of the hot path into oracle/_ref/ (git-ignored; it travels to the GPU box with the
snapshot like a built .so).  TEST / BASELINE INFRASTRUCTURE ONLY.

The reference is pure Python: there is nothing to compile.  Staging the five modules of the path plus the three prior .mat
files (utils_data.py:450-452, 554-557 read them relative to the CWD) lets ``bench.py --impl reference`` and the ``cpu_baseline``
leg time the UNMODIFIED reference (through oracle/shim.py's two stubs) on the GPU box's host cores, where /root/reference does
not exist.  Nothing here is product code and nothing under iip_uavsal_saliency_b200/ reads oracle/_ref.

    python -m oracle.stage_ref            # no-op when /root/reference is absent
"""
from __future__ import annotations

import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference"
DST = os.path.join(HERE, "_ref")
FILES = ("model.py", "model_feature.py", "model_convlstm.py", "utils_data.py", "utils_score_torch.py",
         "gauss_priors.mat", "UAV2_ob_priors_train.mat", "AVS1K_ob_priors_train.mat")


def stage() -> bool:
    if not os.path.isfile(os.path.join(SRC, "model.py")):
        return os.path.isfile(os.path.join(DST, "model.py"))
    os.makedirs(DST, exist_ok=True)
    for f in FILES:
        s, d = os.path.join(SRC, f), os.path.join(DST, f)
        if not os.path.exists(d) or os.path.getmtime(s) > os.path.getmtime(d) or os.path.getsize(s) != os.path.getsize(d):
            shutil.copyfile(s, d)
    return True


if __name__ == "__main__":
    print("staged" if stage() else "reference not available", DST)
