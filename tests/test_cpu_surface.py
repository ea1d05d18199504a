"""CPU-side tests: C-ABI exports, drop-in surface (signatures, state dict, errors), plan structure, packing
arithmetic, MAT v7.3 I/O, prior loaders, sharding logic.  No GPU, no compute calls into the library."""
import inspect
import json
import os
import re

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from iip_uavsal_saliency_b200 import _ext, engine, mat73
from iip_uavsal_saliency_b200 import model as M
from iip_uavsal_saliency_b200 import model_convlstm as MC
from iip_uavsal_saliency_b200 import model_feature as MF
from iip_uavsal_saliency_b200 import utils_data as UD
from iip_uavsal_saliency_b200 import utils_score_torch as US
from oracle import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "uavsal_b200.h")).read()
    declared = set(re.findall(r"\b(uavsal_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    lib = _ext.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.uavsal_version() == 1 and lib.uavsal_arch() == b"sm_100a"
    assert declared == set(_ext.EXPORTS) | {"uavsal_version", "uavsal_arch", "uavsal_last_error"}


def test_state_dict_layout_matches_reference_fixture():
    keys = json.load(open(os.path.join(ROOT, "tests", "golden", "state_dict_keys.json")))
    m = M.UAVSal()
    sd = m.state_dict()
    assert list(sorted(sd)) == sorted(k for k, _, _ in keys) and len(sd) == 685
    for k, shape, dtype in keys:
        assert list(sd[k].shape) == shape and str(sd[k].dtype) == dtype, k
    # a reference-shaped state dict loads strictly, buffers round-trip (quirk Q6: unused features.18, int64 counters)
    new = synth.make_state_dict("lively", 3)
    new["sfnet.features.features.18.1.num_batches_tracked"] = torch.tensor(7)
    m.load_state_dict(new, strict=True)
    back = m.state_dict()
    assert back["sfnet.features.features.18.1.num_batches_tracked"].item() == 7
    assert torch.equal(back["rnn.cell_list.0.rnn_conv.weight"], new["rnn.cell_list.0.rnn_conv.weight"])
    assert sum(v.numel() for v in back.values() if v.dtype == torch.float32) == 13407338 + 117332 - 114


def test_constructor_signatures_and_defaults():
    def defaults(fn):
        return {k: v.default for k, v in inspect.signature(fn).parameters.items() if v.default is not inspect._empty}
    assert defaults(M.UAVSal.__init__) == dict(cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1],
                                               iosize=[360, 640, 45, 80], planes=256, pre_model_path="")
    assert defaults(M.UAVSAL_LSTM.__init__) == dict(cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1],
                                                    iosize=[360, 640, 45, 80], planes=256, pre_model_path="")
    lstm = M.UAVSAL_LSTM()
    assert tuple(lstm.state_dict()["rnn.cell_list.0.rnn_conv.weight"].shape) == (1024, 512, 3, 3) and len(lstm.state_dict()) == 685
    assert defaults(M.dwBlock.__init__) == dict(kernel_size=3, stride=1, expand_ratio=6, dilation=1, res_connect=None)
    assert defaults(M.BasicConv2d.__init__) == dict(kernel_size=3, stride=1, dilation=1, groups=1)
    assert defaults(M.teConv_sub.__init__) == dict(planes=256, time_dims=8, reduction=8, res_connect=False)
    assert defaults(M.STBlock.__init__) == dict(planes=256, time_dims=8, fu_type="sum", res_connect=True)
    assert defaults(M.uavsal_srfnet_aspp.__init__) == dict(cnn_type="mobilenet_v2", planes=[64, 64, 128, 256], last_channel=256)
    assert defaults(MC.ConvLSTM.__init__) == dict(batch_first=False, bias=True, return_all_layers=False)
    assert defaults(MC.ConvTWA.__init__) == dict(batch_first=False, bias=True, return_all_layers=False)
    assert list(inspect.signature(M.UAVSal.forward).parameters) == ["self", "x", "cb", "in_state"]
    assert list(inspect.signature(MC.ConvTWA.forward).parameters) == ["self", "input_tensor", "hidden_state"]
    assert list(inspect.signature(MC.ConvLSTMCell.forward).parameters) == ["self", "input_tensor", "cur_state"]
    for name in ("metric_cc", "metric_nss", "metric_kl", "metric_sim"):
        assert list(inspect.signature(getattr(US, name)).parameters) == ["y_pred", "y_true"]
    # the reference's `metrics` dict (utils_score_torch.py:221-229), AUC metrics included
    assert US.EPS == 2.2204e-16 and set(US.metrics) == {"AUC_shuffled", "AUC_Judd", "AUC_Borji", "NSS", "CC", "SIM", "KLD"}
    assert list(inspect.signature(US.metric_auc_j).parameters) == ["y_pred", "y_true", "jitter"]
    assert list(inspect.signature(US.metric_auc_b).parameters) == ["y_pred", "y_true"]
    assert list(inspect.signature(US.metric_auc_s).parameters) == ["y_pred", "y_true", "shuff_map"]
    for fn in (US.evalscores_vid_torch, US.evalscores_vid_torch_sum):                          # utils_score_torch.py:368, 473
        assert list(inspect.signature(fn).parameters) == ["RootDir", "SalDir", "DataSet", "MethodNames", "keys_order", "batch_size"]


def test_error_behaviour_follows_reference():
    with pytest.raises(ValueError):
        MF.ReMobileNetV2("resnet50")                                   # model_feature.py:54-55
    with pytest.raises(ValueError):
        MC.ConvLSTM((4, 4), 8, 8, 3, 1)                                # kernel_size must be tuple (:227-230)
    with pytest.raises(ValueError):
        MC.ConvTWA((4, 4), 8, [8, 8], (3, 3), 1)                       # inconsistent list length (:143-144)
    with pytest.raises(AssertionError):
        M.dwBlock(8, 8, stride=3)                                      # model.py:78
    m = M.dwBlock(16, 16)
    with pytest.raises(RuntimeError, match="no CPU"):
        m(torch.zeros(1, 16, 8, 8))                                    # product path never falls back to CPU
    with pytest.raises(RuntimeError, match="no CPU"):
        US.metric_cc(torch.zeros(1, 1, 8, 8), torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="CUDA"):
        M.UAVSal().forward(torch.zeros(5, 3, 64, 64), [torch.zeros(5, 8, 8, 8), torch.zeros(5, 20, 8, 8)], None)


def test_plan_structure_of_one_call():
    m = M.UAVSal().eval()
    plan = engine.Plan("cpu", 3, "tc")
    m.build_plan(plan, 20, 360, 640, x_kind=1, post_hw=(360, 640))
    names = [o.name for o in plan.ops]
    # 77 used pointwise convs, the 1536->1 one is the dot; the two stride-2 high-resolution dwBlocks (features.2, features.4) run
    # expand + depthwise fused by default, all eight dwBlocks with <= 32 input channels when asked to
    # ... and the seven large stride-1 blocks whose project conv has 64..256 outputs (st0/st1.sp, fust, gauss1, ob1, fucb, fucbst)
    # run depthwise + project fused
    # ... plus features.1 (32 -> 16, no expand conv), whose depthwise + project run as one fp32 FFMA kernel behind the same entry
    # ... and the twelve narrow stride-1 blocks (cin <= 64, cout <= 64: features.3/5/6/8/9/10, both ST blocks' te.sub, gauss0/1, ob0/1;
    # hidden widths that are not multiples of 64 - 144, 48, 120 - are zero-padded) run as ONE kernel each (uavsal_mbconv_fused)
    assert names.count("uavsal_mbconv_fused") == 12
    assert names.count("uavsal_expand_dw3x3") == 2 and names.count("uavsal_dw_project") == 7 - 2 and names.count("uavsal_dw_project32_hw") == 1
    # ... and the readout's depthwise conv is folded into its 1-output project (dw3x3_dot_sigmoid)
    assert names.count("uavsal_pw_gemm") == 76 - 2 - 8 - 16 - 6 and names.count("uavsal_dw3x3") == 34 - 2 - 8 - 1 - 7 - 3
    pu = engine.Plan("cpu", 3, "tc")
    pu.fuse_mbconv = False
    m.build_plan(pu, 20, 360, 640, x_kind=1, post_hw=(360, 640))
    nu = [o.name for o in pu.ops]
    assert nu.count("uavsal_mbconv_fused") == 0 and nu.count("uavsal_dw_project") == 7 and nu.count("uavsal_pw_gemm") == 76 - 2 - 8
    # the hidden tensors of the widest blocks (>= 1152 channels: st0/st1.sp, fust, cxt0, fucbst, the readout and the three dilated ASPP
    # branches at 12x20; not fucb, whose 64-output dw_project is bound by its depthwise stage) travel as 16-bit fixed-point rows: the expand GEMM writes them
    # (F_OUT_Q16), dw_project / dw3x3 / the readout's dw+dot read them
    assert names.count("uavsal_dw3x3_dot_sigmoid_q16") == 1 and names.count("uavsal_dw3x3_dot_sigmoid") == 0 and names.count("uavsal_dot_sigmoid") == 0
    q16_out = [o for o in plan.ops if o.name == "uavsal_pw_gemm" and o.args[9] & engine.F_OUT_Q16]
    assert len(q16_out) == 9 and all(o.args[9] == engine.F_OUT_Q16 | engine.F_RELU6 and o.args[7] >= engine.Q16_HIDDEN_MIN for o in q16_out)
    assert sum(1 for o in plan.ops if o.name == "uavsal_dw_project" and o.args[12] & engine.F_HID_Q16) == 4
    assert sum(1 for o in plan.ops if o.name == "uavsal_dw3x3" and o.args[1] == engine.PLANE_Q16) == 4          # cxt0 (stride 2) + aspp2..4 (dilation 6 / 12 / 18)
    pq = engine.Plan("cpu", 3, "tc")
    pq.hidden_q16 = False                                               # everything in fp32 rows
    m.build_plan(pq, 20, 360, 640, x_kind=1, post_hw=(360, 640))
    assert [o.name for o in pq.ops].count("uavsal_dw3x3_dot_sigmoid") == 1
    assert not any(o.name == "uavsal_pw_gemm" and o.args[9] & engine.F_OUT_Q16 for o in pq.ops)
    assert not any(o.name == "uavsal_dw3x3" and o.args[1] == engine.PLANE_Q16 for o in pq.ops)
    pf = engine.Plan("cpu", 3, "tc")
    pf.fuse_expand_dw = True
    pf.fuse_dw_project = False
    pf.fuse_mbconv = False
    m.build_plan(pf, 20, 360, 640, x_kind=1, post_hw=(360, 640))
    assert [o.name for o in pf.ops].count("uavsal_expand_dw3x3") == 8
    ps = engine.Plan("cpu", 3, "simt")                                 # the SIMT cross-check engine keeps every conv separate
    m.build_plan(ps, 20, 360, 640, x_kind=1, post_hw=(360, 640))
    ns = [o.name for o in ps.ops]
    assert ns.count("uavsal_pw_gemm_simt") == 76 and ns.count("uavsal_dw3x3") == 34 and ns.count("uavsal_expand_dw3x3") == 0
    assert plan.split == names.index("uavsal_twa_sequence") - 1      # front | back boundary sits before the state pack
    assert names.count("uavsal_conv3x3") == 1 and names.count("uavsal_twa_sequence") == 1
    assert names.count("uavsal_bilinear_ac") == 3 and names.count("uavsal_tdiff_cat") == 2
    assert plan.named["out"].shape == (20, 1, 45, 80) and plan.named["out_u8"].shape == (20, 360, 640)
    assert plan.named["h_out"].shape == (1, 256, 45, 80)
    with pytest.raises(ValueError):
        m.build_plan(engine.Plan("cpu"), 7, 360, 640)                  # n must be a multiple of time_dims
    p2 = engine.Plan("cpu", 3, "tc")
    m.build_plan(p2, 5, 288, 512)
    assert p2.named["out"].shape == (5, 1, 36, 64)


def test_weight_packing_arithmetic():
    torch.manual_seed(0)
    w = torch.randn(24, 20)
    pk = engine.pack_pw_tc(w, 24)
    assert pk.shape == (2, 24, 24) and pk.dtype == torch.bfloat16
    rec = pk[0].float() + pk[1].float()
    assert torch.all(rec[:, 20:] == 0)
    assert (rec[:, :20] - w).abs().max() <= w.abs().max() * 2 ** -16
    assert torch.equal(engine.pack_pw_simt(w, 24)[:20], w.t())
    # BN folding == conv followed by eval-mode batch norm
    blk = M.BasicConv2d(6, 10, 1)
    blk[1].running_mean.normal_(); blk[1].running_var.uniform_(0.5, 2); blk[1].weight.data.normal_(); blk[1].bias.data.normal_()
    blk.eval()
    x = torch.randn(2, 6, 5, 5)
    wf, bf = blk.folded()
    ref = blk[1](blk[0](x))
    mine = F.conv2d(x, wf, bf)
    assert (ref - mine).abs().max() < 1e-5
    # implicit-GEMM K ordering: k = (ky*3+kx)*Cin + ci
    w4 = torch.randn(7, 5, 3, 3)
    x = torch.randn(1, 5, 6, 6)
    cols = F.unfold(x, 3, padding=1).reshape(5, 9, 36).permute(1, 0, 2).reshape(45, 36)      # (tap, ci) major
    assert (engine.conv3x3_as_2d(w4) @ cols - F.conv2d(x, w4, padding=1).reshape(7, 36)).abs().max() < 1e-5
    # LSTM gate interleave: packed row 4*c+g <- reference row g*ch+c
    ch = 3
    wl = torch.arange(4 * ch).float().reshape(4 * ch, 1)
    il = engine.interleave_gates(wl, ch).reshape(-1)
    for c in range(ch):
        for g in range(4):
            assert il[4 * c + g].item() == g * ch + c
    assert torch.equal(engine.pack_dw(torch.arange(18.).reshape(2, 1, 3, 3))[:, 1], torch.arange(9., 18.))


def test_mat73_reads_reference_prior_format_and_roundtrips(tmp_path, gold_dir):
    pri = np.load(os.path.join(gold_dir, "priors.npz"))
    path = str(tmp_path / "gauss_priors.mat")
    mat73.savemat(path, {"PriorMaps": pri["gauss"]})
    back = mat73.loadmat(path)["PriorMaps"]
    assert back.dtype == np.float32 and np.array_equal(back, pri["gauss"])
    sal = np.random.RandomState(0).randint(0, 256, (36, 64, 1, 7)).astype(np.uint8)
    mat73.savemat(str(tmp_path / "s.mat"), {"salmap": sal})
    assert np.array_equal(mat73.loadmat(str(tmp_path / "s.mat"))["salmap"], sal)
    ref = "/root/reference/gauss_priors.mat"
    if os.path.exists(ref):       # authoring container only: the real file (chunked + shuffle + deflate + fletcher32)
        assert np.array_equal(mat73.loadmat(ref)["PriorMaps"], pri["gauss"])


def test_prior_loaders_and_host_helpers(tmp_path, gold_dir, monkeypatch):
    pri = np.load(os.path.join(gold_dir, "priors.npz"))
    monkeypatch.chdir(tmp_path)
    g = UD.get_guasspriors(3, 45, 80, 8)                               # no .mat in CWD -> closed form (utils_data.py:453-457)
    assert g.shape == (3, 45, 80, 8) and np.array_equal(g[1], pri["gauss"])
    mat73.savemat("gauss_priors.mat", {"PriorMaps": pri["gauss"]})
    mat73.savemat("UAV2_ob_priors_train.mat", {"PriorMaps": pri["uav2_u8"].astype(np.float32) / 255})
    assert np.array_equal(UD.get_guasspriors(2)[0], pri["gauss"])
    assert UD.get_ob_priors("", "UAV2", "train", 2).shape == (2, 45, 80, 20)
    assert UD.get_guasspriors(1, 36, 64).max() == 0 and UD.get_ob_priors("", "UAV2", "train", 1, 36, 64).dtype == np.uint8   # quirk Q4
    with pytest.raises(NotImplementedError):
        UD.read_ob_priors("", "UAV2", "test")
    post = np.load(os.path.join(gold_dir, "post_u8.npz"))
    assert np.array_equal(UD.normalize_data(post["x8"]), post["nd"])
    with pytest.raises(ValueError):
        UD.normalize_data(np.zeros((4, 4), np.uint8))
    assert np.array_equal(UD.np2mat(np.array([0.5, 1.5, 2.5, 300., -3.])), np.array([0, 2, 2, 255, 0], np.uint8))   # rint = half-to-even


def test_eval_host_helpers_match_reference(tmp_path):
    """Host-side helpers of the evaluation driver (utils_score_torch.py:248-357): the dataset's summed fixation map at the
    native size and through resize_fixation, against the unmodified reference on the same synthetic tree
    (tests/golden/eval_driver.npz); getALLFix_vid / getshufmap shapes and value ranges (their effect on the scores is pinned
    by the AUC_shuffled column of the driver test)."""
    import numpy as np
    from oracle import synth
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "eval_driver.npz"))
    root, sal = str(tmp_path) + "/data/", str(tmp_path) + "/res/"
    synth.make_eval_dataset(root, sal, 0)
    assert np.array_equal(US.getSumFix_vid(root + "fixations/maps/", "UAV2", size=(36, 64)), g["sumfix_native"])
    assert np.array_equal(US.getSumFix_vid(root + "fixations/maps/", "UAV2", size=(45, 80)), g["sumfix_resized"])
    pts = US.getALLFix_vid(root + "fixations/maps/", "UAV2")
    assert len(pts) == 10 and all(p.shape == (12, 2) and p.min() >= 0 and p.max() < 1 for p in pts)
    np.random.seed(0)
    sm = US.getshufmap(pts, size=(36, 64))
    assert sm.shape == (36, 64) and sm.dtype == np.uint8 and 12 <= sm.sum() <= 120
    assert pts[0].max() < 1                                          # the list entries are not scaled in place


def test_two_pass_arena_plan_shares_memory_only_between_disjoint_lifetimes():
    """Plan.build's measure pass + layout solver (engine.py), run here on the CPU: every allocation gets an offset in one arena;
    two buffers overlap in memory only if one's last op precedes the other's first; module-boundary tensors and everything
    reachable from ``named`` (debug taps) are persistent; the second pass reproduces the op list of a direct plan."""
    from iip_uavsal_saliency_b200 import engine
    from iip_uavsal_saliency_b200.model import UAVSal
    torch.manual_seed(0)
    m = UAVSal(iosize=[288, 512, 36, 64]).eval()
    build = lambda plan: m.build_plan(plan, 10, 288, 512, x_kind=1, post_hw=(288, 512), taps=True)
    probe = engine.Plan("cpu", 3, "tc", mode="measure")
    build(probe)
    offs, total = probe.solve_layout()
    allocs = probe._allocs
    trans = [a for a in allocs if not a.pinned]
    assert trans and all(a.first is not None and a.last >= a.first for a in trans)
    for i, a in enumerate(allocs):
        for b in allocs[i + 1:]:
            mem_overlap = not (a.off + a.nbytes <= b.off or b.off + b.nbytes <= a.off)
            if a.pinned or b.pinned:
                assert not mem_overlap
            elif not (a.last < b.first or b.last < a.first):
                assert not mem_overlap, (a.idx, b.idx)
    taps = probe.named["taps"]
    assert all(buf.root.pinned for buf, _, _ in taps.values())          # debug taps must survive the run
    plan = engine.Plan("cpu", 3, "tc", mode="arena", layout=(offs, total))
    build(plan)
    direct = engine.Plan("cpu", 3, "tc")
    build(direct)
    assert [o.name for o in plan.ops] == [o.name for o in direct.ops] == [o.name for o in probe.ops]
    assert len(plan._allocs) == len(allocs) and plan.arena_bytes == total < direct.arena_bytes / 3
    lo, hi = plan._arena.data_ptr(), plan._arena.data_ptr() + total
    for name in ("x_in", "out", "out_u8", "h_in", "h_out"):
        t = plan.named[name]
        assert lo <= t.data_ptr() and t.data_ptr() + t.numel() * t.element_size() <= hi and t.shape == direct.named[name].shape
    sf, _, _ = plan.named["taps"]["sfnet"]
    assert sf.t.shape == (2, 10 * 36 * 64, 256) and sf.to_float().shape == (10 * 36 * 64, 256)
    # packed weights are cached on the owning conv and follow the parameters' versions
    conv = m.fust_layer[0].conv[0][0]
    cache = conv.__dict__["_uavsal_packed"]
    (key, (sig, wt, b)), = [(k, v) for k, v in cache.items() if k[0] == engine.W_ROWS_SPLIT]
    assert wt.shape == (2, 1536, 256) and b.shape == (1536,)
    again = engine.Plan("cpu", 3, "tc")
    wt2, _ = again.packed(m.fust_layer[0].conv[0].wspec(), engine.W_ROWS_SPLIT, 1536, 256)
    assert wt2 is wt
    with torch.no_grad():
        conv.weight.mul_(2.0)
    wt3, _ = again.packed(m.fust_layer[0].conv[0].wspec(), engine.W_ROWS_SPLIT, 1536, 256)
    assert wt3 is not wt and torch.equal(wt3[0].float(), (wt[0].float() * 2))
    m.invalidate()
    assert "_uavsal_packed" not in conv.__dict__


def test_pack_reference_matches_the_module_level_folding():
    """W.pack_reference (the torch restatement uavsal_pack_weights is checked against on the GPU) equals the explicit
    fold_bn / conv3x3_as_2d / interleave_gates / split_bf16 pipeline it replaced."""
    from iip_uavsal_saliency_b200 import engine
    from iip_uavsal_saliency_b200.blocks import BasicConv2d
    torch.manual_seed(1)
    c = BasicConv2d(24, 40, 3)
    c[1].running_mean.normal_(); c[1].running_var.uniform_(0.5, 2.0); c[1].weight.data.uniform_(0.5, 1.5); c[1].bias.data.normal_()
    wf, bf = c.folded()
    wp, b = c.wspec().pack_reference(engine.W_ROWS_SPLIT, 48, 9 * 24 + 8)
    ref = torch.zeros(48, 9 * 24 + 8)
    ref[:40, :216] = engine.conv3x3_as_2d(wf)
    assert torch.equal(wp, engine.split_bf16(ref)) and torch.equal(b[:40], bf) and not b[40:].any()
    wc, _ = c.wspec().pack_reference(engine.W_COLS_F32, 40, 216)
    assert torch.equal(wc, engine.conv3x3_as_2d(wf).t())
    w4 = torch.randn(32, 12, 3, 3)
    bias = torch.randn(32)
    wl, bl = engine.W(w4, bias=bias).pack_reference(engine.W_ROWS_F32, 32, 108, gates=4)
    assert torch.equal(wl, engine.conv3x3_as_2d(engine.interleave_gates(w4, 8))) and torch.equal(bl, engine.interleave_gates(bias, 8))
    dw = BasicConv2d(16, 16, 3, groups=16)
    wd, bd = dw.wspec().pack_reference(engine.W_COLS_F32, 16, 9)
    assert torch.equal(wd, engine.pack_dw(dw.folded()[0])) and torch.equal(bd, dw.folded()[1])


def test_backbone_wrappers_surface():
    """ReResNet / ReVGG (model_feature.py:72-128): error behaviour of the name checks, torchvision's state-dict layout (without
    the classification head the reference discards), the reference's feature_loader / feature_inplanes tables (model.py:14-33),
    and the plan structure of a ResNet-50 UAVSal."""
    import torchvision
    with pytest.raises(ValueError):
        MF.ReResNet("mobilenet_v2")                      # not a resnet name (:75-76)
    with pytest.raises(NotImplementedError):
        MF.ReResNet("resnext50_32x4d")                   # a torchvision resnet the reference has no loader for (:77-78)
    with pytest.raises(ValueError):
        MF.ReVGG("resnet18")
    with pytest.raises(NotImplementedError):
        MF.ReVGG("vgg11")
    assert sorted(M.feature_loader) == sorted(["vgg16", "resnet18", "resnet34", "resnet50", "resnet101", "resnet152", "mobilenet_v2"])
    assert M.feature_inplanes["resnet50"] == [256, 512, 1024, 2048] and M.feature_inplanes["vgg16"] == [128, 256, 512, 512]
    for name in ("resnet18", "resnet34", "resnet50", "vgg16"):
        mine = (MF.ReVGG if name == "vgg16" else MF.ReResNet)(name)
        tv = getattr(torchvision.models, name)(weights=None)
        ref = {("features." + k): v for k, v in tv.features.state_dict().items()} if name == "vgg16" else \
            {k: v for k, v in tv.state_dict().items() if not k.startswith("fc.")}
        got = mine.state_dict()
        assert list(got) == list(ref), name
        assert all(got[k].shape == ref[k].shape and got[k].dtype == ref[k].dtype for k in ref), name
    m = M.UAVSal(cnn_type="resnet50", iosize=[96, 160, 12, 20])
    plan = engine.Plan("cpu", 3, "tc")
    m.build_plan(plan, 5, 96, 160, x_kind=1)
    names = [o.name for o in plan.ops]
    assert names[0] == "uavsal_conv_first" and names[1] == "uavsal_maxpool"
    assert names.count("uavsal_add_act") == 16 and names.count("uavsal_conv3x3") == 16 + 1        # 16 bottlenecks + conv_last
    assert names.count("uavsal_maxpool") == 1 + 2 * 3                                           # stem pool + (conv2, shortcut) subsampling of layer2-4
    assert plan.named["out"].shape == (5, 1, 12, 20)


def test_plan_cache_keys_and_eviction(monkeypatch):
    """_kernel_module: plans are keyed on one (data_ptr, _version) pair per parameter / buffer (a tracked in-place update or a
    re-allocation of ANY tensor gives a new plan), evicted least-recently-used by arena bytes, and dropped by invalidate()."""
    from iip_uavsal_saliency_b200.blocks import dwBlock
    torch.manual_seed(0)
    blk = dwBlock(16, 16).eval()
    built = []

    def builder(plan):
        built.append(plan)
        plan.alloc(1000, 64)                                       # 256 000 bytes of arena

    dev = torch.device("cpu")
    p1 = blk._cached_plan((dev, "a"), builder)
    assert blk._cached_plan((dev, "a"), builder) is p1 and len(built) == 1
    sig = blk._weights_signature()
    assert len(sig) == len(list(blk.parameters())) + len(list(blk.buffers())) and all(len(t) == 2 for t in sig)
    with torch.no_grad():
        blk.conv[1][1].running_var.mul_(1.5)                        # a tracked in-place update of one buffer
    assert blk._weights_signature() != sig
    p2 = blk._cached_plan((dev, "a"), builder)
    assert p2 is not p1 and len(built) == 2
    # swapping two equally shaped parameters' storage changes the signature (the XOR of pointers the first version used did not)
    a, b = blk.conv[0][1].weight, blk.conv[0][1].bias
    sig = blk._weights_signature()
    a.data, b.data = b.data, a.data
    assert blk._weights_signature() != sig
    # eviction by bytes: with a 0 GB budget every new plan displaces the others
    monkeypatch.setenv("UAVSAL_PLAN_CACHE_GB", "0")
    blk._cached_plan((dev, "b"), builder)
    blk._cached_plan((dev, "c"), builder)
    assert len(blk._plan_cache()) == 1
    blk.invalidate()
    assert len(blk._plan_cache()) == 0


def test_hidden_row_format_rules_and_q16_view(monkeypatch):
    """Plan.hidden_fmt: which storage the tensor between a dwBlock's expand conv and its depthwise conv gets (model.py:90-92), and the
    host-side view of q16 rows (q * 6 / 65535, the value the kernels read back)."""
    p = engine.Plan("cpu", 3, "tc")
    assert p.hidden_fmt(1536) == engine.FMT_Q16 and p.hidden_fmt(1152) == engine.FMT_Q16 and p.hidden_fmt(960) == engine.FMT_F32
    assert p.hidden_fmt(1920, 6, 12 * 20) == engine.FMT_Q16           # dilated: whole-image kernel, small maps only
    assert p.hidden_fmt(1920, 6, 45 * 80) == engine.FMT_SPLIT and p.hidden_fmt(960, 6, 240) == engine.FMT_SPLIT
    p.hidden_q16 = False
    assert p.hidden_fmt(1536) == engine.FMT_F32 and p.hidden_fmt(1920, 6, 240) == engine.FMT_SPLIT
    assert engine.Plan("cpu", 3, "simt").hidden_fmt(1536) == engine.FMT_SPLIT      # the cross-check engines keep split-bf16 everywhere
    monkeypatch.setenv("UAVSAL_HIDDEN_Q16", "0")
    assert engine.Plan("cpu", 3, "tc").hidden_fmt(1536) == engine.FMT_F32
    monkeypatch.delenv("UAVSAL_HIDDEN_Q16")
    b = engine.Plan("cpu", 3, "tc").alloc_q16(4, 16)
    assert b.q16 and b.plain and not b.f32 and b.plane == engine.PLANE_Q16 and b.t.dtype == torch.int16
    q = torch.tensor([0, 1, 32767, 32768, 65535, 10922], dtype=torch.int32)
    b.t[0, :6] = q.to(torch.int16)                                     # the same 16 bits
    v = b.to_float()[0, :6]
    assert torch.equal(v, q.float() * np.float32(6.0 / 65535.0)) and v[4].item() == pytest.approx(6.0, abs=1e-6) and v[0].item() == 0.0
    s = b.slot(8, 8)
    assert s.ptr == b.ptr + 16 and s.fmt == engine.FMT_Q16            # 2 bytes per element
