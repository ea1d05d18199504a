"""Reference checkpoint loader: whole pickled modules (Demo_Train_Test.py:160,174 / Demo_Test.py:39) are read
without the reference's classes being importable.  Fixtures: oracle/make_ckpt_fixture.py (reference modules pickled in the
authoring container, in the zip and the legacy stream format; the same script round-trips the full 685-key UAVSal there)."""
import io
import os
import pickle

import numpy as np
import pytest
import torch

from iip_uavsal_saliency_b200 import checkpoint

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", ["ckpt_small_zip.pth", "ckpt_small_legacy.pth", "ckpt_small_statedict.pth"])
def test_reference_pickles_load_without_reference_classes(name):
    exp = np.load(os.path.join(GOLD, "ckpt_expected.npz"))
    sd = checkpoint.load_reference_state_dict(os.path.join(GOLD, name))
    assert list(sd.keys()) == list(exp.files) and len(sd) == 73
    for k in exp.files:
        assert sd[k].dtype == torch.from_numpy(exp[k]).dtype and np.array_equal(sd[k].numpy(), exp[k]), k
    # reference (model.*, model_convlstm.*) and legacy torchvision (torchvision.models.mobilenet.*) layers are all in there
    assert "block.conv.1.0.weight" in sd and "rnn.cell_list.0.rnn_conv.weight" in sd and "tv.1.conv.0.0.weight" in sd
    assert sd["tv.0.1.num_batches_tracked"].dtype == torch.int64


def test_subtree_loads_into_this_packages_modules():
    """The pickled reference dwBlock / ConvTWA weights load strictly into this package's classes of the same name."""
    from iip_uavsal_saliency_b200.model import dwBlock
    from iip_uavsal_saliency_b200.model_convlstm import ConvTWA
    sd = checkpoint.load_reference_state_dict(os.path.join(GOLD, "ckpt_small_zip.pth"))
    blk = dwBlock(8, 8, expand_ratio=6)
    r = blk.load_state_dict({k[len("block."):]: v for k, v in sd.items() if k.startswith("block.")}, strict=True)
    assert not r.missing_keys and not r.unexpected_keys
    rnn = ConvTWA((6, 8), 8, 8, (3, 3), 1, batch_first=True, bias=False)
    r = rnn.load_state_dict({k[len("rnn."):]: v for k, v in sd.items() if k.startswith("rnn.")}, strict=True)
    assert not r.missing_keys and not r.unexpected_keys


def test_wrappers_and_prefixes(tmp_path):
    sd0 = checkpoint.load_reference_state_dict(os.path.join(GOLD, "ckpt_small_statedict.pth"))
    p = tmp_path / "wrapped.pth"
    torch.save({"state_dict": {"module." + k: v for k, v in sd0.items()}, "epoch": 3}, p)
    sd = checkpoint.load_reference_state_dict(str(p))
    assert list(sd.keys()) == list(sd0.keys())


class _Evil:
    def __reduce__(self):
        return (os.system, ("echo pwned",))


def test_foreign_callables_are_refused(tmp_path):
    p = tmp_path / "evil.pth"
    with open(p, "wb") as fh:
        torch.save({"w": torch.zeros(1), "x": _Evil()}, fh)
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_reference_state_dict(str(p))
    q = tmp_path / "garbage.pth"
    q.write_bytes(b"not a checkpoint")
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_reference_state_dict(str(q))
    r = tmp_path / "list.pth"
    torch.save([1, 2, 3], r)
    with pytest.raises(checkpoint.CheckpointError):
        checkpoint.load_reference_state_dict(str(r))


def test_unpickler_refuses_code_execution_gadgets(tmp_path):
    """The loader resolves only an explicit list of globals: dotted names (protocol 4 walks attributes: ("torch", "os.system")),
    builtins.getattr / eval, and any module outside the list are refused before anything is called."""
    import pickle
    import pickletools  # noqa: F401
    from iip_uavsal_saliency_b200 import checkpoint as ck

    def payload(mod, name, proto):
        if proto >= 4:      # STACK_GLOBAL resolves dotted names
            return (b"\x80\x04" + b"\x8c" + bytes([len(mod)]) + mod.encode() + b"\x8c" + bytes([len(name)]) + name.encode() + b"\x93"
                    + b"\x8c\x04true" + b"\x85R.")
        return b"\x80\x02c" + mod.encode() + b"\n" + name.encode() + b"\nU\x04true\x85R."

    for mod, name in (("torch", "os.system"), ("builtins", "getattr"), ("builtins", "eval"), ("os", "system"), ("torch.hub", "load"),
                      ("torch.storage", "_load_from_bytes"), ("numpy", "load"), ("subprocess", "Popen")):
        for proto in (2, 4):
            with pytest.raises(ck.CheckpointError):
                ck._Unpickler(__import__("io").BytesIO(payload(mod, name, proto))).load()
    # and through the public entry point: a zip-less legacy file that is just such a pickle
    bad = tmp_path / "evil.pth"
    bad.write_bytes(payload("torch", "os.system", 4))
    with pytest.raises(ck.CheckpointError):
        ck.load_reference_state_dict(str(bad))
