"""Make ``from model import *; from utils_score_torch import *`` (Demo_Test.py:6-10) resolve to this package."""
import importlib
import sys

_NAMES = ("model", "model_feature", "model_convlstm", "utils_data", "utils_score_torch")


def install(force: bool = False):
    for name in _NAMES:
        if name in sys.modules and not force:
            continue
        sys.modules[name] = importlib.import_module("iip_uavsal_saliency_b200." + name)
