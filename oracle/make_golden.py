"""Regenerate tests/golden/*.npz by running the REAL reference (unmodified, from /root/reference) through
oracle/shim.py.  TEST INFRASTRUCTURE ONLY.  Runs only in the authoring container:

    python -m oracle.make_golden [--only NAME]

Fixtures (all inputs come from oracle/synth.py seeds, so tests can rebuild them without the reference):
  plumbing_288.npz     BASELINE config #1: 16 frames 288x512, batch_size=1,time_dims=5 → 15 maps; stock+lively
  clip64_360.npz       BASELINE config #2: 64 frames 360x640, Demo_Test grouping 20/20/20 → 60 maps, lively
  call20_trace.npz     one 20-frame call at 360x640 (B=4,T=5; quirks Q2,Q3) with sampled per-stage traces
  metrics_pairs.npz    CC/NSS/KLD/SIM of the reference on 8 synthetic pairs + known-answer cases
  rnn_small.npz        ConvLSTM / ConvTWA on small shapes
  post_u8.npz          postprocess_predictions + np2mat (real cv2.resize) on random maps
  priors.npz           PriorMaps of the three .mat files (ob priors as exact uint8*255 where they are k/255)
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import shim, synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TRACE_SAMPLES = 4096


def sample_idx(numel: int, key: str) -> np.ndarray:
    rs = np.random.RandomState(abs(hash_str(key)) % (2 ** 31))
    return rs.randint(0, numel, size=min(TRACE_SAMPLES, numel))


def hash_str(s: str) -> int:
    h = 2166136261
    for ch in s.encode():
        h = ((h ^ ch) * 16777619) & 0xFFFFFFFF
    return h


def ref_model(ref, iosize, sd, time_dims=5):
    m = ref.model.UAVSal(cnn_type="mobilenet_v2", time_dims=time_dims, num_stblock=2, bias_type=[1, 1, 1],
                         iosize=iosize, planes=256, pre_model_path="").eval()
    m.load_state_dict(sd, strict=True)
    return m


def run_demo_loop(ref, model, frames_u8, gauss_nchw1, ob_nchw1, time_dims, batch_size, out_hw):
    """Demo_Test.test's inner loop (Demo_Test.py:68-91) driving the reference's own functions."""
    ud = ref.utils_data
    F_ = frames_u8.shape[0]
    count_bs = F_ // time_dims
    keep = count_bs * time_dims
    vid = frames_u8[:keep].transpose((0, 3, 1, 2))
    per_call = batch_size * time_dims
    h, w = gauss_nchw1.shape[2:]
    state = [torch.zeros(1, 256, h, w)]
    maps, u8 = [], []
    with torch.no_grad():
        for i in range(int(np.ceil(count_bs / batch_size))):
            x = vid[i * per_call:(i + 1) * per_call]
            x = torch.tensor(ud.normalize_data(x)).float()
            n = x.shape[0]
            cb = [torch.tensor(np.repeat(gauss_nchw1, n, 0)).float(), torch.tensor(np.repeat(ob_nchw1, n, 0)).float()]
            out, st = model(x, cb, state)
            state = [st[0].detach()]
            o = out.data.cpu().numpy()
            maps.append(o)
            for j in range(n):
                u8.append(ud.np2mat(ud.postprocess_predictions(o[j, 0, :, :], out_hw[0], out_hw[1])))
    return np.concatenate(maps, 0), np.stack(u8, 0), state[0].numpy()


def uav2_priors(ref):
    with shim.reference_cwd():
        g = ref.utils_data.get_guasspriors(1, 45, 80, 8).transpose((0, 3, 1, 2))
        o = ref.utils_data.get_ob_priors("", "UAV2", "train", 1, 45, 80).transpose((0, 3, 1, 2))
    return np.ascontiguousarray(g), np.ascontiguousarray(o)


def gen_priors(ref):
    from iip_uavsal_saliency_b200 import mat73
    g = mat73.loadmat(os.path.join(shim.REFERENCE_ROOT, "gauss_priors.mat"))["PriorMaps"]
    u = mat73.loadmat(os.path.join(shim.REFERENCE_ROOT, "UAV2_ob_priors_train.mat"))["PriorMaps"]
    a = mat73.loadmat(os.path.join(shim.REFERENCE_ROOT, "AVS1K_ob_priors_train.mat"))["PriorMaps"]
    u8 = np.rint(u * 255.0).astype(np.uint8)
    assert np.array_equal(u8.astype(np.float32) / 255, u), "UAV2 priors are not exactly k/255"
    np.savez_compressed(os.path.join(GOLD, "priors.npz"), gauss=g, uav2_u8=u8, avs1k=a.astype(np.float16),
                        avs1k_sample_idx=sample_idx(a.size, "avs1k"), avs1k_sample=a.ravel()[sample_idx(a.size, "avs1k")])


def gen_backbones(ref):
    """backbones.npz: UAVSal on the alternative backbones (model_feature.ReResNet / ReVGG, model.py:14-33) - one 5-frame call at
    96x160 through the unmodified reference; weights from synth.make_state_dict_like over the reference model's own key table."""
    out = {}
    clip = synth.make_clip(7, 5, 96, 160)
    x = torch.tensor(ref.utils_data.normalize_data(clip.transpose(0, 3, 1, 2))).float()
    g, o = synth.make_priors(5, 12, 20, seed=3)
    cb = [torch.from_numpy(g), torch.from_numpy(o)]
    for cnn in ("resnet18", "resnet50", "vgg16"):
        m = ref.model.UAVSal(cnn_type=cnn, time_dims=5, num_stblock=2, bias_type=[1, 1, 1], iosize=[96, 160, 12, 20], planes=256, pre_model_path="").eval()
        sd = synth.make_state_dict_like(synth.key_table_of(m), 11)
        m.load_state_dict(sd, strict=True)
        with torch.no_grad():
            levels = m.sfnet.features(x)
            sf = m.sfnet(x)
            y, st = m(x, cb, [torch.zeros(1, 256, 12, 20)])
        out[cnn + "_keys"] = np.array(len(sd))
        for i, lv in enumerate(levels):
            a = lv.numpy()
            out["%s_level%d_shape" % (cnn, i)] = np.array(a.shape)
            out["%s_level%d" % (cnn, i)] = a.ravel()[sample_idx(a.size, "%s_level%d" % (cnn, i))]
        a = sf.numpy()
        out[cnn + "_sfnet"] = a.ravel()[sample_idx(a.size, cnn + "_sfnet")]
        out[cnn + "_out"] = y.numpy()
        h = st[0].numpy()
        out[cnn + "_h"] = h.ravel()[sample_idx(h.size, cnn + "_h")]
        print(cnn, "out range", float(y.min()), float(y.max()), "levels", [tuple(l.shape) for l in levels])
    np.savez_compressed(os.path.join(GOLD, "backbones.npz"), **out)


def gen_plumbing(ref):
    out = {}
    clip = synth.make_clip(0, 16, 288, 512)
    g, o = synth.make_priors(1, 36, 64, seed=0)
    for kind in ("stock", "lively"):
        sd = synth.make_state_dict(kind, 0)
        m = ref_model(ref, [288, 512, 36, 64], sd)
        maps, u8, h = run_demo_loop(ref, m, clip, g, o, 5, 1, (288, 512))
        out[kind + "_maps"] = maps
        out[kind + "_u8_frames"] = u8[[0, 7, 14]]
        out[kind + "_h_last_sample"] = h.ravel()[sample_idx(h.size, "h_last")]
    # the reference's own prior loaders at 36x64 give all-zero priors (quirk Q4)
    with shim.reference_cwd():
        gz = ref.utils_data.get_guasspriors(1, 36, 64, 8)
        oz = ref.utils_data.get_ob_priors("", "UAV2", "train", 1, 36, 64)
    out["q4_gauss_max"] = np.array(gz.max())
    out["q4_ob_max"] = np.array(oz.max())
    np.savez_compressed(os.path.join(GOLD, "plumbing_288.npz"), **out)


def gen_clip64(ref):
    clip = synth.make_clip(2, 64, 360, 640)
    g, o = uav2_priors(ref)
    sd = synth.make_state_dict("lively", 0)
    m = ref_model(ref, [360, 640, 45, 80], sd)
    maps, u8, h = run_demo_loop(ref, m, clip, g, o, 5, 4, (360, 640))
    np.savez_compressed(os.path.join(GOLD, "clip64_360.npz"), maps=maps, u8_frames=u8[[0, 19, 20, 59]],
                        u8_frame_idx=np.array([0, 19, 20, 59]),
                        h_last_sample=h.ravel()[sample_idx(h.size, "h_last")])


def gen_call20_trace(ref):
    clip = synth.make_clip(1, 20, 360, 640)
    g, o = uav2_priors(ref)
    sd = synth.make_state_dict("lively", 0)
    m = ref_model(ref, [360, 640, 45, 80], sd)
    ud = ref.utils_data
    x = torch.tensor(ud.normalize_data(clip.transpose((0, 3, 1, 2)))).float()
    cb = [torch.tensor(np.repeat(g, 20, 0)).float(), torch.tensor(np.repeat(o, 20, 0)).float()]
    rs = np.random.RandomState(7)
    h0 = torch.tensor(rs.randn(1, 256, 45, 80).astype(np.float32) * 0.5)
    taps = {
        "c3": "sfnet.features.features.6", "c4": "sfnet.features.features.13", "c5": "sfnet.features.features.17",
        "sfnet": "sfnet", "st_layer.0": "st_layer.0", "st_layer.1": "st_layer.1", "fust": "fust_layer",
        "cb_gauss": "gauss_cb_layer", "cb_ob": "ob_cb_layer", "fucb": "fucb_layer", "fucbst": "fucbst_layer",
    }
    got = {}
    mods = dict(m.named_modules())
    hooks = []
    for name, path in taps.items():
        hooks.append(mods[path].register_forward_hook(
            lambda mod, inp, outp, name=name: got.__setitem__(name, outp.detach().numpy())))
    with torch.no_grad():
        out, st = m(x, cb, [h0])
    for hk in hooks:
        hk.remove()
    res = {"out": out.numpy(), "h_last_sample": st[0].numpy().ravel()[sample_idx(st[0].numel(), "h_last")]}
    for name, arr in got.items():
        idx = sample_idx(arr.size, name)
        res["trace_" + name] = arr.ravel()[idx]
        res["shape_" + name] = np.array(arr.shape)
    np.savez_compressed(os.path.join(GOLD, "call20_trace.npz"), **res)


def gen_metrics(ref):
    us = ref.utils_score_torch
    pred, true = synth.make_metric_pairs(8, 360, 640, seed=0)
    p, t = torch.from_numpy(pred), torch.from_numpy(true)
    vals = torch.cat([us.metric_cc(p, t), us.metric_nss(p, t), us.metric_kl(p, t), us.metric_sim(p, t)], 1).numpy()
    # known answers: identical maps, all-zero prediction, small random
    rs = np.random.RandomState(11)
    tt = torch.from_numpy(rs.rand(2, 2, 8, 8).astype(np.float32))
    same = tt[:, 0:1].clone()
    zero = torch.zeros(2, 1, 8, 8)
    ka_same = torch.cat([us.metric_cc(same, tt), us.metric_nss(same, tt), us.metric_kl(same, tt), us.metric_sim(same, tt)], 1).numpy()
    ka_zero = torch.cat([us.metric_cc(zero, tt), us.metric_nss(zero, tt), us.metric_kl(zero, tt), us.metric_sim(zero, tt)], 1).numpy()
    np.savez_compressed(os.path.join(GOLD, "metrics_pairs.npz"), values=vals, ka_true=tt.numpy(), ka_same=ka_same,
                        ka_zero=ka_zero)


def gen_auc(ref):
    """AUC-Judd / Borji / shuffled of the UNMODIFIED reference on seeded inputs (global torch / numpy generators seeded
    right before each call, which is what makes the reference's Monte-Carlo draws reproducible)."""
    us = ref.utils_score_torch
    pred, true, shuf = synth.make_auc_case(0)
    p, t, o = torch.from_numpy(pred), torch.from_numpy(true), torch.from_numpy(shuf)
    res = {}
    res["judd_nojitter"] = us.metric_auc_j(p, t, jitter=0).numpy()
    torch.manual_seed(1234)
    res["judd_jitter"] = us.metric_auc_j(p, t).numpy()
    np.random.seed(4321)
    res["borji"] = us.metric_auc_b(p, t).numpy()
    np.random.seed(987)
    res["shuffled"] = us.metric_auc_s(p, t, o).numpy()
    np.savez_compressed(os.path.join(GOLD, "auc_metrics.npz"), **res)
    print({k: v.ravel().round(5).tolist() for k, v in res.items()})


def gen_frontend(ref):
    """utils_data.padding (:321-343) and preprocess_videos (:255-287) of the UNMODIFIED reference: seeded frames through
    padding() for the wide / tall / exact-2x / identity geometries, and a tiny MJPG clip (committed next to the values)
    through the whole preprocess_videos (decode + letterbox + BGR->RGB, normalize=False and True)."""
    import cv2
    ud = ref.utils_data
    rs = np.random.RandomState(21)
    res = {}
    for name, (sh, sw, r, c) in {"wide": (54, 160, 72, 128), "tall": (150, 100, 72, 128), "x2": (144, 256, 72, 128),
                                 "same": (72, 128, 72, 128), "up": (30, 40, 72, 128)}.items():
        img = rs.randint(0, 256, (sh, sw, 3)).astype(np.uint8)
        res["pad_in_" + name] = img
        res["pad_out_" + name] = ud.padding(img, r, c, 3)
    path = os.path.join(GOLD, "clip_tiny.avi")
    wr = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 25.0, (200, 120))
    clip = synth.make_clip(9, 6, 120, 200)                      # (6,120,200,3) uint8
    for f in clip:
        wr.write(np.ascontiguousarray(f[:, :, ::-1]))
    wr.release()
    ims, nframes, height, width = ud.preprocess_videos(path, 72, 128, normalize=False)
    res["vid_u8"], res["vid_meta"] = ims, np.array([nframes, height, width])
    imsn, _, _, _ = ud.preprocess_videos(path, 72, 128, frames=4, normalize=True)
    res["vid_norm4"] = imsn
    np.savez_compressed(os.path.join(GOLD, "frontend.npz"), **res)
    print("frontend:", ims.shape, ims.dtype, imsn.shape, imsn.dtype, os.path.getsize(path), "B avi")


def gen_lstm_model(ref):
    """UAVSAL_LSTM (model.py:960-1076, the Table-V ablation) of the unmodified reference: two chained calls of 10 frames at
    288x512 (state [h, c] handed over as ConvLSTM.forward expects it: one pair per layer)."""
    clip = synth.make_clip(4, 20, 288, 512)
    g, o = synth.make_priors(1, 36, 64, seed=0)
    m = ref.model.UAVSAL_LSTM(cnn_type="mobilenet_v2", time_dims=5, num_stblock=2, bias_type=[1, 1, 1], iosize=[288, 512, 36, 64],
                              planes=256, pre_model_path="").eval()
    m.load_state_dict(synth.make_state_dict_lstm(0), strict=True)
    x = torch.tensor(ref.utils_data.normalize_data(clip.transpose((0, 3, 1, 2)))).float()
    cb = [torch.tensor(np.repeat(g, 10, 0)).float(), torch.tensor(np.repeat(o, 10, 0)).float()]
    state = [[torch.zeros(1, 256, 36, 64), torch.zeros(1, 256, 36, 64)]]
    res = {}
    with torch.no_grad():
        for call in range(2):
            out, st = m(x[call * 10:(call + 1) * 10], cb, state)
            res["out%d" % call] = out.numpy()
            res["h%d" % call] = st[0].numpy().ravel()[sample_idx(st[0].numel(), "lstm_h")]
            res["c%d" % call] = st[1].numpy().ravel()[sample_idx(st[1].numel(), "lstm_c")]
            state = [st]
    np.savez_compressed(os.path.join(GOLD, "uavsal_lstm_288.npz"), **res)
    print("uavsal_lstm:", res["out1"].shape, float(res["out1"].min()), float(res["out1"].max()), float(np.abs(res["c1"]).max()))


def gen_eval_driver(ref):
    """evalscores_vid_torch (utils_score_torch.py:473-582) of the unmodified reference on synth.make_eval_dataset, all seven
    metrics, generators seeded.  The reference spells np.int / np.NaN (:347, :570), which numpy 2 dropped: the two names are
    aliased for the run (no source change)."""
    import tempfile
    us = ref.utils_score_torch
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "NaN"):
        np.NaN = np.nan
    with tempfile.TemporaryDirectory() as td:
        root, sal = td + "/data/", td + "/res/"
        synth.make_eval_dataset(root, sal, 0)
        np.random.seed(11)
        torch.manual_seed(11)
        us.evalscores_vid_torch(root, sal, "UAV2", ["UAVSal"], batch_size=3)
        from iip_uavsal_saliency_b200 import mat73
        res = {n: mat73.loadmat(sal + "Scores/UAVSal/Score_%s.mat" % n)["iscore"] for n in ("vidA", "vidB")}
        res["sumfix_native"] = us.getSumFix_vid(root + "fixations/maps/", "UAV2", size=(36, 64))
        res["sumfix_resized"] = us.getSumFix_vid(root + "fixations/maps/", "UAV2", size=(45, 80))      # resize_fixation path (:248-263)
    np.savez_compressed(os.path.join(GOLD, "eval_driver.npz"), keys=np.array(list(us.keys_order)), **res)
    print(list(us.keys_order)); print(res["vidA"].round(4)); print(res["vidB"].round(4))


def gen_eval_driver_sum(ref):
    """evalscores_vid_torch_sum (utils_score_torch.py:368-470) of the unmodified reference.  Its `shuffle_map != []` test (:424)
    raises under numpy 2 ("operands could not be broadcast"); numpy < 1.25 answered True for an array.  That one legacy answer is
    restored without touching the source: the hdf5storage stub hands ShufMap over as an ndarray subclass whose __ne__ says True
    to an empty list.  The shuffle map file is written beforehand with the reference's own getSumFix_vid (the branch that
    computes it in place would compare a plain ndarray)."""
    import tempfile
    import hdf5storage as h5
    from iip_uavsal_saliency_b200 import mat73
    us = ref.utils_score_torch
    if not hasattr(np, "int"):
        np.int = int
    if not hasattr(np, "NaN"):
        np.NaN = np.nan

    class LegacyNe(np.ndarray):
        def __ne__(self, other):
            if isinstance(other, list) and len(other) == 0:
                return True
            return np.ndarray.__ne__(self, other)

    orig = h5.loadmat
    h5.loadmat = lambda path, *a, **k: {kk: (v.view(LegacyNe) if kk == "ShufMap" else v) for kk, v in orig(path).items()}
    try:
        with tempfile.TemporaryDirectory() as td:
            root, sal = td + "/data/", td + "/res/"
            synth.make_eval_dataset(root, sal, 0, halve_second=False)
            sm = us.getSumFix_vid(root + "fixations/maps/", "UAV2")
            mat73.savemat(root + "Shuffle_UAV2.mat", {"ShufMap": sm})
            np.random.seed(12)
            torch.manual_seed(12)
            us.evalscores_vid_torch_sum(root, sal, "UAV2", ["UAVSal"], batch_size=3)
            res = {n: mat73.loadmat(sal + "Scores_sum/UAVSal/Score_%s.mat" % n)["iscore"] for n in ("vidA", "vidB")}
            res["shufmap_sum"] = np.array(sm.sum())
    finally:
        h5.loadmat = orig
    np.savez_compressed(os.path.join(GOLD, "eval_driver_sum.npz"), keys=np.array(list(us.keys_order)), **res)
    print(res["vidA"].round(4)); print(res["vidB"].round(4))


def gen_demo_test(ref):
    """Demo_Test.test (Demo_Test.py:30-95) of the unmodified reference, end to end on the committed MJPG clip: decode, letterbox
    to 360x640, one 5-frame call with the real UAV2 priors, post-process to the video's size, salmap .mat.  The model file is
    the reference's own pickle of a UAVSal holding the 'lively' weights."""
    import shutil
    import tempfile
    sys.path.insert(0, shim.REFERENCE_ROOT)
    try:
        import Demo_Test as dt
    finally:
        sys.path.remove(shim.REFERENCE_ROOT)
    dt.DataSet_Train = "UAV2"                                      # a module global set under __main__ (Demo_Test.py:121)
    # the reference's torch.load(model_path) (Demo_Test.py:39) predates torch 2.6's weights_only=True default: restore the old
    # default through torch's own environment switch (no source change)
    os.environ["TORCH_FORCE_NO_WEIGHTS_ONLY_LOAD"] = "1"
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(td + "/videos")
        shutil.copy(os.path.join(GOLD, "clip_tiny.avi"), td + "/videos/clip_tiny.avi")
        m = ref_model(ref, [360, 640, 45, 80], synth.make_state_dict("lively", 0))
        torch.save(m, td + "/model.pth")
        # ConvTWACell.init_hidden ends in .cuda() (model_convlstm.py:295, quirk Q5) and test() starts from x_state = None: on this
        # GPU-less container Tensor.cuda is made a no-op for the duration of the run
        real_cuda = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
        try:
            with shim.reference_cwd():
                dt.test(td + "/videos/", td + "/out/", td + "/model.pth", iosize=[360, 640, 45, 80], batch_size=4, time_dims=5)
        finally:
            torch.Tensor.cuda = real_cuda
        from iip_uavsal_saliency_b200 import mat73
        sal = mat73.loadmat(td + "/out/UAVSal/clip_tiny.mat")["salmap"]
    np.savez_compressed(os.path.join(GOLD, "demo_test.npz"), salmap=sal)
    print("demo_test:", sal.shape, sal.dtype, int(sal.max()), float(sal.mean()))


def gen_rnn_small(ref):
    mc = ref.model_convlstm
    res = {}
    rs = np.random.RandomState(5)
    for bias in (False, True):
        torch.manual_seed(3)
        net = mc.ConvLSTM((10, 12), 8, 16, (3, 3), 1, batch_first=True, bias=bias).eval()
        w = rs.randn(64, 24, 3, 3).astype(np.float32) * 0.2
        net.cell_list[0].rnn_conv.weight.data.copy_(torch.from_numpy(w))
        if bias:
            b = rs.randn(64).astype(np.float32) * 0.3
            net.cell_list[0].rnn_conv.bias.data.copy_(torch.from_numpy(b))
            res["lstm_b"] = b
        x = rs.randn(2, 3, 8, 10, 12).astype(np.float32)
        h0 = rs.randn(2, 16, 10, 12).astype(np.float32) * 0.5
        c0 = rs.randn(2, 16, 10, 12).astype(np.float32) * 0.5
        with torch.no_grad():
            y, (h, c) = net(torch.from_numpy(x), [[torch.from_numpy(h0), torch.from_numpy(c0)]])
        tag = "lstm_bias" if bias else "lstm"
        res.update({tag + "_w": w, tag + "_x": x, tag + "_h0": h0, tag + "_c0": c0, tag + "_y": y.numpy(),
                    tag + "_h": h.numpy(), tag + "_c": c.numpy()})
    net = mc.ConvTWA((10, 12), 16, 16, (3, 3), 1, batch_first=True, bias=False).eval()
    w = rs.randn(16, 32, 3, 3).astype(np.float32) * 0.2
    net.cell_list[0].rnn_conv.weight.data.copy_(torch.from_numpy(w))
    x = rs.randn(1, 4, 16, 10, 12).astype(np.float32)
    h0 = rs.randn(1, 16, 10, 12).astype(np.float32)
    with torch.no_grad():
        y, st = net(torch.from_numpy(x), [torch.from_numpy(h0)])
    res.update({"twa_w": w, "twa_x": x, "twa_h0": h0, "twa_y": y.numpy(), "twa_h": st[0].numpy()})
    np.savez_compressed(os.path.join(GOLD, "rnn_small.npz"), **res)


def gen_post(ref):
    ud = ref.utils_data
    rs = np.random.RandomState(9)
    smooth = synth._upsample_linear(rs.rand(6, 10), 45, 80).astype(np.float32)
    m1 = (0.2 + 0.7 * smooth + 0.02 * rs.rand(45, 80)).astype(np.float32)
    u1 = ud.np2mat(ud.postprocess_predictions(m1.copy(), 360, 640))
    u2 = ud.np2mat(ud.postprocess_predictions(m1.copy(), 720, 1280))
    m3 = rs.rand(36, 64).astype(np.float32)
    u3 = ud.np2mat(ud.postprocess_predictions(m3.copy(), 300, 500))     # rows_rate > cols_rate → crop columns
    u4 = ud.np2mat(ud.postprocess_predictions(m3.copy(), 270, 512))     # crop rows
    x8 = rs.randint(0, 256, size=(2, 3, 16, 20)).astype(np.uint8)
    nd = ud.normalize_data(x8)
    np.savez_compressed(os.path.join(GOLD, "post_u8.npz"), m1=m1, u1=u1, u2=u2, m3=m3, u3=u3, u4=u4, x8=x8, nd=nd)


GENS = {"priors": gen_priors, "plumbing": gen_plumbing, "clip64": gen_clip64, "call20": gen_call20_trace,
        "metrics": gen_metrics, "auc": gen_auc, "frontend": gen_frontend, "lstm_model": gen_lstm_model, "eval": gen_eval_driver, "eval_sum": gen_eval_driver_sum, "demo": gen_demo_test, "rnn": gen_rnn_small, "post": gen_post, "backbones": gen_backbones}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count() or 1)
    ref = shim.load()
    os.makedirs(GOLD, exist_ok=True)
    for name, fn in GENS.items():
        if a.only and name not in a.only.split(","):
            continue
        print("golden:", name, flush=True)
        fn(ref)


if __name__ == "__main__":
    main()
