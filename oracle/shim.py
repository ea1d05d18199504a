"""Import the UNMODIFIED reference modules from /root/reference, offline.  TEST INFRASTRUCTURE ONLY.

Two stubs are needed:
  1. ``hdf5storage`` is imported at module top by utils_data.py:6 and utils_score_torch.py:9 but is not
     installed; a stub module backed by ``iip_uavsal_saliency_b200.mat73`` is injected.
  2. ``ReMobileNetV2.__init__`` calls ``mobilenet_v2(pretrained=True)`` (model_feature.py:59), which needs
     the network; ``model_feature.feature_loader['mobilenet_v2']`` is replaced by a constructor that builds
     the same torchvision architecture with ``weights=None``.

The reference reads its prior ``.mat`` files relative to the CWD (utils_data.py:450-452, 554-557); use
``reference_cwd()`` around calls to its prior loaders.  The modules come from /root/reference (the authoring container) or,
where that does not exist (the GPU box), from the copy oracle/stage_ref.py staged under oracle/_ref; GPU-box TESTS use the
committed fixtures, only bench.py's CPU-baseline legs run the staged reference.
"""
from __future__ import annotations

import contextlib
import importlib
import os
import sys
import types

_REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_STAGED = os.path.join(_REPO_ROOT, "oracle", "_ref")          # oracle/stage_ref.py: the same files, staged for the GPU box


def _pick_root() -> str:
    env = os.environ.get("UAVSAL_REFERENCE_ROOT")
    if env:
        return env
    return "/root/reference" if os.path.isfile("/root/reference/model.py") else _STAGED


REFERENCE_ROOT = _pick_root()


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


def _install_hdf5storage_stub():
    if "hdf5storage" in sys.modules:
        return
    if _REPO_ROOT not in sys.path:
        sys.path.insert(0, _REPO_ROOT)
    from iip_uavsal_saliency_b200 import mat73

    stub = types.ModuleType("hdf5storage")
    stub.loadmat = lambda path, *a, **k: mat73.loadmat(path)
    stub.savemat = lambda path, d, *a, **k: mat73.savemat(path, d)
    sys.modules["hdf5storage"] = stub


_REF_MODULE_NAMES = ("model", "model_feature", "model_convlstm", "utils_data", "utils_score_torch")


def load():
    """Return a namespace with the reference modules: .model .model_feature .model_convlstm .utils_data
    .utils_score_torch.  They are imported under their own top-level names, as the reference expects."""
    if not available():
        raise RuntimeError("reference checkout not present at %s" % REFERENCE_ROOT)
    _install_hdf5storage_stub()
    for name in _REF_MODULE_NAMES:
        mod = sys.modules.get(name)
        if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
            raise RuntimeError("module name %r already taken by %s" % (name, mod.__file__))
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        mods = {n: importlib.import_module(n) for n in _REF_MODULE_NAMES}
    finally:
        sys.path.remove(REFERENCE_ROOT)
    import torchvision

    # every backbone constructor of model_feature.feature_loader downloads ImageNet weights (pretrained=True, :59, :80, :115)
    for name in list(mods["model_feature"].feature_loader):
        mods["model_feature"].feature_loader[name] = (lambda ctor: (lambda pretrained=True: ctor(weights=None)))(getattr(torchvision.models, name))
    return types.SimpleNamespace(**mods)


@contextlib.contextmanager
def reference_cwd():
    old = os.getcwd()
    os.chdir(REFERENCE_ROOT)
    try:
        yield
    finally:
        os.chdir(old)
