"""Read the reference's published / trained checkpoints without the reference's source tree.

The reference saves WHOLE pickled modules (``torch.save(model, path)``, /root/reference/Demo_Train_Test.py:160,174) and
loads them back with ``torch.load(model_path).state_dict()`` (/root/reference/Demo_Test.py:39, model.py:339).  Unpickling
such a file needs every class it mentions to be importable under its original name: ``model.UAVSal``, ``model.dwBlock``,
``model_convlstm.ConvTWA`` ..., and — for files written with the torchvision versions the reference pins (0.5.0 / 0.8.2,
README.md:27,34) — ``torchvision.models.mobilenet.ConvBNReLU / InvertedResidual / MobileNetV2``, which current
torchvision no longer has.

``load_reference_state_dict`` therefore unpickles with a remapping ``find_class``: every class of the reference's own
modules and of ``torchvision`` is replaced by an empty ``nn.Module`` stand-in that only carries ``_parameters / _buffers /
_modules`` — which is all ``state_dict()`` walks.  Everything else must be on an explicit list of (module, name) pairs:
the tensor / storage / parameter rebuild helpers, ``torch.Size`` / dtypes / storage classes, ``collections.OrderedDict``,
classes below ``torch.nn.modules`` that ARE ``nn.Module`` subclasses, numpy's array reconstructors and a handful of plain
builtins.  Dotted names (pickle protocol 4 resolves ``os.system`` through ``torch`` that way), ``builtins.getattr`` / ``eval``
and any other callable are refused, so a hostile file cannot reach arbitrary code through this loader (the reference's own
``torch.load`` executes whatever the pickle says).  The result is the plain 685-key state dict of SURVEY App. A, ready for
``UAVSal.load_state_dict(..., strict=True)``.

Files that already hold a state dict (``OrderedDict`` of tensors), or a dict with a ``"state_dict"`` entry, are accepted too.
"""
from __future__ import annotations

import collections
import io
import pickle
import types
import warnings
from typing import Dict

import torch
from torch import nn

# modules whose classes are replaced by parameter-carrying stand-ins
_STUB_ROOTS = ("model", "model_feature", "model_convlstm", "torchvision", "__main__")
# exact (module, name) pairs that resolve normally
_EXACT = {
    ("collections", "OrderedDict"), ("collections", "defaultdict"),
    ("torch._utils", "_rebuild_tensor"), ("torch._utils", "_rebuild_tensor_v2"), ("torch._utils", "_rebuild_parameter"),
    ("torch._utils", "_rebuild_parameter_with_state"), ("torch._utils", "_rebuild_qtensor"),
    ("torch", "Size"), ("torch", "device"), ("torch.serialization", "_get_layout"), ("torch.nn.parameter", "Parameter"),
    ("torch._tensor", "_rebuild_from_type_v2"), ("torch", "Tensor"),
    ("numpy.core.multiarray", "_reconstruct"), ("numpy._core.multiarray", "_reconstruct"), ("numpy.core.multiarray", "scalar"),
    ("numpy._core.multiarray", "scalar"), ("numpy", "ndarray"), ("numpy", "dtype"),
    ("_codecs", "encode"), ("copyreg", "_reconstructor"),
}
_BUILTINS_OK = {"set", "frozenset", "list", "dict", "tuple", "int", "float", "bool", "str", "bytes", "slice", "range", "complex",
                "object", "bytearray"}


class CheckpointError(ValueError):
    pass


class _Stub(nn.Module):
    """Stand-in for a reference / torchvision module class: holds whatever ``__dict__`` the pickle restores."""

    def forward(self, *a, **k):  # pragma: no cover - never called
        raise RuntimeError("checkpoint stand-in modules cannot be executed; use their state_dict()")


_stub_cache: Dict[str, type] = {}


def _stub_for(mod_name: str, name: str) -> type:
    key = mod_name + "." + name
    cls = _stub_cache.get(key)
    if cls is None:
        cls = type(name, (_Stub,), {"__module__": mod_name})
        _stub_cache[key] = cls
    return cls


class _Unpickler(pickle.Unpickler):
    def find_class(self, mod_name, name):
        if not name.isidentifier():
            # protocol 4 resolves dotted names attribute by attribute: ("torch", "os.system") would be arbitrary code
            raise CheckpointError("checkpoint refers to the dotted name %s.%s" % (mod_name, name))
        root = mod_name.split(".", 1)[0]
        if root in _STUB_ROOTS:
            return _stub_for(mod_name, name)
        if mod_name in ("__builtin__", "builtins"):   # protocol-2 streams (torch.save's default) use the Python 2 module name
            if name not in _BUILTINS_OK:
                raise CheckpointError("checkpoint refers to builtins.%s, which a model file has no business doing" % name)
            return super().find_class("builtins", name)
        if (mod_name, name) in _EXACT:
            return super().find_class(mod_name, name)
        if mod_name == "torch":                        # dtypes and storage classes
            obj = getattr(torch, name, None)
            if isinstance(obj, torch.dtype) or (isinstance(obj, type) and name.endswith("Storage")):
                return obj
        if mod_name == "torch.storage" and name in ("UntypedStorage", "TypedStorage"):
            return super().find_class(mod_name, name)
        if mod_name == "torch.nn.modules" or mod_name.startswith("torch.nn.modules."):
            obj = super().find_class(mod_name, name)
            if isinstance(obj, type) and issubclass(obj, nn.Module):
                return obj
        raise CheckpointError("checkpoint refers to %s.%s, which is neither a tensor / stock-layer object nor a reference model class" % (mod_name, name))


def _pickle_module() -> types.ModuleType:
    """A ``pickle``-shaped module for ``torch.load(pickle_module=...)`` whose Unpickler remaps classes."""
    m = types.ModuleType("uavsal_checkpoint_pickle")
    m.Unpickler = _Unpickler
    m.Pickler = pickle.Pickler
    m.load = lambda f, **kw: _Unpickler(f, **kw).load()
    m.loads = lambda b, **kw: _Unpickler(io.BytesIO(b), **kw).load()
    m.dump, m.dumps = pickle.dump, pickle.dumps
    m.HIGHEST_PROTOCOL, m.DEFAULT_PROTOCOL = pickle.HIGHEST_PROTOCOL, pickle.DEFAULT_PROTOCOL
    m.PickleError, m.UnpicklingError = pickle.PickleError, pickle.UnpicklingError
    return m


def _to_state_dict(obj) -> "collections.OrderedDict[str, torch.Tensor]":
    if isinstance(obj, nn.Module):
        # checkpoints written by torch 1.4 / 1.7 predate attributes that today's state_dict()/BatchNorm code reads
        for m in obj.modules():
            d = m.__dict__
            d.setdefault("_non_persistent_buffers_set", set())
            d.setdefault("_state_dict_hooks", collections.OrderedDict())
            d.setdefault("_state_dict_pre_hooks", collections.OrderedDict())
            d.setdefault("_load_state_dict_pre_hooks", collections.OrderedDict())
            d.setdefault("_load_state_dict_post_hooks", collections.OrderedDict())
            d.setdefault("_backward_hooks", collections.OrderedDict())
            d.setdefault("_backward_pre_hooks", collections.OrderedDict())
            d.setdefault("_forward_hooks", collections.OrderedDict())
            d.setdefault("_forward_pre_hooks", collections.OrderedDict())
        sd = obj.state_dict()
    elif isinstance(obj, dict) and "state_dict" in obj and isinstance(obj["state_dict"], dict):
        sd = obj["state_dict"]
    elif isinstance(obj, dict):
        sd = obj
    else:
        raise CheckpointError("unsupported checkpoint payload of type %s" % type(obj).__name__)
    out = collections.OrderedDict()
    for k, v in sd.items():
        if not isinstance(k, str) or not isinstance(v, torch.Tensor):
            raise CheckpointError("checkpoint entry %r is not a named tensor" % (k,))
        if k.startswith("module."):            # nn.DataParallel wrappers
            k = k[len("module."):]
        out[k] = v.detach()
    return out


def load_reference_state_dict(path, map_location="cpu") -> "collections.OrderedDict[str, torch.Tensor]":
    """State dict of a reference checkpoint: a whole pickled module (``torch.save(model)``), a state dict, or a
    ``{"state_dict": ...}`` wrapper.  The reference's code does not have to be importable."""
    try:
        with warnings.catch_warnings():        # the legacy stream format carries class sources torch wants to diff: not ours to check
            warnings.filterwarnings("ignore", message="Couldn't retrieve source code for container")
            obj = torch.load(path, map_location=map_location, pickle_module=_pickle_module(), weights_only=False)
    except CheckpointError:
        raise
    except (pickle.UnpicklingError, AttributeError, ModuleNotFoundError, RuntimeError, EOFError) as e:
        raise CheckpointError("cannot read checkpoint %s: %s" % (path, e)) from e
    return _to_state_dict(obj)


def load_into(model: nn.Module, path, strict: bool = True, map_location="cpu"):
    """``model.load_state_dict(torch.load(path).state_dict())`` (Demo_Test.py:39) for a model of THIS package."""
    sd = load_reference_state_dict(path, map_location=map_location)
    return model.load_state_dict(sd, strict=strict)
