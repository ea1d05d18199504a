"""Generate the checkpoint-loader fixtures (tests/golden/ckpt_*.pth + ckpt_expected.npz).  TEST INFRASTRUCTURE ONLY.

Runs only in the authoring container (needs /root/reference through oracle/shim.py):

    python oracle/make_ckpt_fixture.py

1. Whole-model check (nothing committed - the file is 54 MB): the UNMODIFIED reference ``UAVSal`` is pickled the way
   Demo_Train_Test.py:160 does (``torch.save(model, path)``), in the zip format and in the legacy (torch <= 1.5) stream format,
   and read back by ``iip_uavsal_saliency_b200.checkpoint`` in a fresh interpreter that cannot import the reference.
2. Small committed fixtures: a composite of reference modules (``model.dwBlock``, ``model.teConv_sub``,
   ``model_convlstm.ConvTWA``) and of stand-ins named like the torchvision 0.5 classes the published files contain
   (``torchvision.models.mobilenet.ConvBNReLU / InvertedResidual``), pickled in both formats, plus the expected tensors.
"""
import os
import subprocess
import sys
import tempfile
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from torch import nn

from oracle import shim

GOLD = os.path.join(ROOT, "tests", "golden")

_CHECK = r"""
import sys, numpy as np, torch
sys.path.insert(0, %(root)r)
from iip_uavsal_saliency_b200 import checkpoint
from iip_uavsal_saliency_b200.model import UAVSal
exp = np.load(%(npz)r)
for p in %(paths)r:
    sd = checkpoint.load_reference_state_dict(p)
    assert list(sd.keys()) == list(exp.files), (len(sd), len(exp.files))
    for k in exp.files:
        assert np.array_equal(sd[k].numpy(), exp[k]), k
    m = UAVSal()
    r = checkpoint.load_into(m, p, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    assert "model_feature" not in sys.modules or "iip_uavsal" in sys.modules["model_feature"].__name__
print("whole-model checkpoint: %%d keys round-trip through both formats, strict load ok" %% len(exp.files))
"""


def legacy_torchvision():
    """Classes named like torchvision 0.5's mobilenet.py (the versions the reference pins), for the pickle stream only."""
    mod = types.ModuleType("torchvision.models.mobilenet")

    class ConvBNReLU(nn.Sequential):
        def __init__(self, i, o, k=3, s=1, groups=1):
            super().__init__(nn.Conv2d(i, o, k, s, (k - 1) // 2, groups=groups, bias=False), nn.BatchNorm2d(o), nn.ReLU6(inplace=True))

    class InvertedResidual(nn.Module):
        def __init__(self, inp, oup, stride, expand_ratio):
            super().__init__()
            hidden = inp * expand_ratio
            self.use_res_connect = stride == 1 and inp == oup
            self.conv = nn.Sequential(ConvBNReLU(inp, hidden, 1), ConvBNReLU(hidden, hidden, 3, stride, hidden),
                                      nn.Conv2d(hidden, oup, 1, 1, 0, bias=False), nn.BatchNorm2d(oup))

    for c in (ConvBNReLU, InvertedResidual):
        c.__module__ = mod.__name__
        c.__qualname__ = c.__name__
        setattr(mod, c.__name__, c)
    return mod


def randomize(m, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for t in list(m.parameters()) + [b for b in m.buffers() if b.dtype.is_floating_point]:
            t.copy_(torch.randn(t.shape, generator=g) * 0.1 + (1.0 if t.ndim == 1 else 0.0))
        for b in m.buffers():
            if not b.dtype.is_floating_point:
                b.fill_(7)


def main():
    ref = shim.load()
    os.makedirs(GOLD, exist_ok=True)
    # ---- 1. whole model, both formats, verified in a fresh interpreter ----
    with shim.reference_cwd():
        model = ref.model.UAVSal(cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1], iosize=[360, 640, 45, 80], planes=256)
    randomize(model, 1)
    with tempfile.TemporaryDirectory() as td:
        pz, pl, npz = os.path.join(td, "uavsal_zip.pth"), os.path.join(td, "uavsal_legacy.pth"), os.path.join(td, "expected.npz")
        torch.save(model, pz)
        torch.save(model, pl, _use_new_zipfile_serialization=False)
        np.savez(npz, **{k: v.numpy() for k, v in model.state_dict().items()})
        out = subprocess.run([sys.executable, "-c", _CHECK % {"root": ROOT, "npz": npz, "paths": [pz, pl]}], capture_output=True, text=True, cwd=td)
        print(out.stdout.strip(), out.stderr.strip()[-2000:])
        assert out.returncode == 0
    # ---- 2. small committed fixtures ----
    tv = legacy_torchvision()
    sys.modules[tv.__name__] = tv
    try:
        small = nn.ModuleDict({
            "block": ref.model.dwBlock(8, 8, expand_ratio=6),
            "te": ref.model.teConv_sub(16, planes=16, time_dims=5, reduction=4),
            "rnn": ref.model_convlstm.ConvTWA((6, 8), 8, 8, (3, 3), 1, batch_first=True, bias=False),
            "tv": nn.Sequential(tv.ConvBNReLU(3, 8, 3, 2), tv.InvertedResidual(8, 8, 1, 6)),
        })
        randomize(small, 2)
        torch.save(small, os.path.join(GOLD, "ckpt_small_zip.pth"))
        torch.save(small, os.path.join(GOLD, "ckpt_small_legacy.pth"), _use_new_zipfile_serialization=False)
        torch.save(small.state_dict(), os.path.join(GOLD, "ckpt_small_statedict.pth"))
        np.savez_compressed(os.path.join(GOLD, "ckpt_expected.npz"), **{k: v.numpy() for k, v in small.state_dict().items()})
        print("small fixtures: %d keys, %s" % (len(small.state_dict()), ", ".join("%s %d B" % (f, os.path.getsize(os.path.join(GOLD, f)))
                                                                                   for f in sorted(os.listdir(GOLD)) if f.startswith("ckpt_"))))
    finally:
        del sys.modules[tv.__name__]


if __name__ == "__main__":
    main()
