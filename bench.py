#!/usr/bin/env python
"""bench.py — UAVSal frames/sec at 360x640 on N x B200 (BASELINE.json metric), one process per GPU.

    python bench.py --gpus 1 --steps K --warmup W                 # product arm (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W         # reference arm: the unmodified reference on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...         # clip-sharded (BASELINE config #4), strong scaling

A "step" = one pass of the hot path over one batch of synthetic input = `--clips` (default 16) clips of 64 uint8 frames at
360x640, split over the ranks as clip c -> rank c % world (BASELINE config #4's partition; 20 steps = 320 clips), every clip
processed with Demo_Test's grouping (batch_size=4, time_dims=5 -> calls of 20/20/20 frames, 60 saliency maps per clip, the 4
tail frames dropped exactly as the reference does).  `value` counts produced maps per second over all ranks with the frames
already resident in HBM; `e2e` is the same loop with host (pinned) uint8 frames in and host uint8 maps out, copies inside
the timed region.  The total work per step is fixed, so `scaling` is "strong"; `weak_scaling` repeats the measurement with a
fixed number of clips per GPU.

The line also carries the two other single-kernel BASELINE configs, measured in the same process:
  convlstm_config3   ConvLSTM gate conv + cell update, hidden 256 ch at 45x80, 64 steps, batch 8 (rank 0)
  metrics_config5    CC/NSS/KLD/SIM on 16 384 synthetic 360x640 pairs sharded over the ranks + the final all-reduce
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES, H, W, MH, MW = 64, 360, 640, 45, 80
BATCH, T = 4, 5
OUT_PER_CLIP = (FRAMES // T) * T
WORKLOAD = ("UAVSal inference, synthetic 64-frame clips at 360x640 (BASELINE config #2 clips, config #4 partition), Demo_Test grouping "
            "batch_size=4 x time_dims=5 -> 60 maps/clip, 'lively' random weights, UAV2-shaped priors")
METRIC = "UAVSal frames/sec at 360x640"


def load_priors():
    import numpy as np
    p = os.path.join(ROOT, "tests", "golden", "priors.npz")
    if os.path.exists(p):
        z = np.load(p)
        return z["gauss"], z["uav2_u8"].astype(np.float32) / 255
    from iip_uavsal_saliency_b200 import synth
    g, o = synth.make_priors(1, MH, MW)
    return g[0].transpose(1, 2, 0), o[0].transpose(1, 2, 0)


def load_peaks():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["bf16_tflops_sustained"]), float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json; tensor: sustained)"
    except Exception:
        return 1590.0 * 0.88, 6650.0, "fallback (B200_PROFILING.md)"


def load_traffic():
    """DRAM bytes per launch measured by ncu in this round (profiles/r02_traffic.json, written by tools/summarize_traffic.py from
    the committed captures; the file names its sources)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "25"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
        rows = [r for t, r in self.rows if self.t0 is None or (self.t0 <= t <= (self.t1 or t) + 0.03)]
        if not rows and self.rows:
            rows = [r for _, r in self.rows[-2:]]
        sm, mx, reasons, pw = [], 0.0, set(), []
        for r in rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(pw) if pw else None}


# ---------------------------------------------------------------------------------------------------
# CPU legs: the UNMODIFIED reference (staged under oracle/_ref by oracle/stage_ref.py, driven through oracle/shim.py) on the
# host cores; the oracle port only where the staged reference is missing
# ---------------------------------------------------------------------------------------------------
class CpuArm:
    def __init__(self):
        import torch
        from iip_uavsal_saliency_b200 import synth
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        self.sd = synth.make_state_dict("lively", 0)
        self.gauss, self.ob = load_priors()
        self.kind = "port"
        self.ref = self.model = None
        try:
            from oracle import shim
            if shim.available():
                self.ref = shim.load()
                self.model = self.ref.model.UAVSal(cnn_type="mobilenet_v2", time_dims=T, num_stblock=2, bias_type=[1, 1, 1],
                                                   iosize=[H, W, MH, MW], planes=256, pre_model_path="").eval()
                self.model.load_state_dict(self.sd, strict=True)
                self.kind = "reference"
        except Exception as e:                                      # pragma: no cover - environment dependent
            print("bench.py: staged reference unavailable (%s); timing the oracle port" % e, file=sys.stderr)
            self.ref = self.model = None

    def uavsal_call(self, clip_u8, n_frames=20):
        """One bounded sample: ONE n-frame call (batch_size x time_dims frames of a clip): normalise, forward, post-process to
        uint8 - Demo_Test.py:77-91.  Returns seconds."""
        import numpy as np
        import torch
        g1 = np.ascontiguousarray(self.gauss.transpose(2, 0, 1)[None])
        o1 = np.ascontiguousarray(self.ob.transpose(2, 0, 1)[None])
        t0 = time.perf_counter()
        if self.kind == "reference":
            from oracle.make_golden import run_demo_loop
            run_demo_loop(self.ref, self.model, clip_u8[:n_frames], g1, o1, T, n_frames // T, (H, W))
        else:
            from oracle import cpu_ref
            x = torch.from_numpy(cpu_ref.normalize_data(clip_u8[:n_frames].transpose(0, 3, 1, 2)))
            cb = [torch.from_numpy(np.repeat(g1, n_frames, 0).copy()), torch.from_numpy(np.repeat(o1, n_frames, 0).copy())]
            out, _ = cpu_ref.uavsal_forward(self.sd, x, cb, torch.zeros(1, 256, MH, MW), time_dims=T)
            o = out.numpy()
            for j in range(n_frames):
                cpu_ref.im2uint8(cpu_ref.postprocess_predictions(o[j, 0], H, W))
        return time.perf_counter() - t0

    def convlstm_steps(self, steps=2):
        """config #3 on the host: `steps` steps of ConvLSTM(256 -> 256, 3x3) at 45x80, batch 8 (reference module, else the port)."""
        import torch
        torch.manual_seed(0)
        x = torch.randn(8, steps, 256, MH, MW)
        h0, c0 = torch.zeros(8, 256, MH, MW), torch.zeros(8, 256, MH, MW)
        with torch.no_grad():
            if self.kind == "reference":
                net = self.ref.model_convlstm.ConvLSTM((MH, MW), 256, 256, (3, 3), 1, batch_first=True, bias=False).eval()
                net(x[:, :1], [[h0, c0]])
                t0 = time.perf_counter()
                net(x, [[h0, c0]])
            else:
                from oracle import cpu_ref
                w = torch.randn(1024, 512, 3, 3) * 0.01
                cpu_ref.lstm_sequence(w, None, x[:, :1], h0, c0)
                t0 = time.perf_counter()
                cpu_ref.lstm_sequence(w, None, x, h0, c0)
        return (time.perf_counter() - t0) / steps

    def metrics_pairs(self, pred, true):
        """config #5 on the host: the four metric functions on a batch of pairs.  Returns seconds."""
        import torch
        p, t = torch.from_numpy(pred), torch.from_numpy(true)
        with torch.no_grad():
            t0 = time.perf_counter()
            if self.kind == "reference":
                us = self.ref.utils_score_torch
                for fn in (us.metric_cc, us.metric_nss, us.metric_kl, us.metric_sim):
                    fn(p, t)                                    # host tensors: the CPU path, whatever the module-level `device` says
            else:
                from oracle import cpu_ref
                cpu_ref.metrics4(p, t)
        return time.perf_counter() - t0


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from iip_uavsal_saliency_b200 import synth
    os.environ.setdefault("CUDA_VISIBLE_DEVICES", "")          # the reference arm is the CPU path: its `device` globals must say cpu
    arm = CpuArm()
    clip = synth.make_clip(100, 20, H, W)
    for _ in range(max(0, min(args.warmup, 1))):
        arm.uavsal_call(clip)
    ts = [arm.uavsal_call(clip) for _ in range(max(1, min(args.steps, 30)))]        # ~2 s per step on 16 cores
    total = sum(ts)
    fps = 20 * len(ts) / total
    sample = ("one 20-frame call (batch_size=4 x time_dims=5) of a 64-frame clip per step, incl. normalisation and the CPU post-process "
              "(Demo_Test.py:77-91); %d steps timed" % len(ts))
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(ts), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(ts), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": sample},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": arm.cores, "kind": arm.kind, "sample": sample},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


# ---------------------------------------------------------------------------------------------------
# product arm
# ---------------------------------------------------------------------------------------------------
def _pw_is_pair(m, k, n, sms=148):
    """Mirror of want_cluster() in csrc/gemm_tc.cu: the GEMMs that run as gemm_tc2_kernel<MODE_PW, EPI_STD, TERMS, CL=2>."""
    n16 = (n + 15) // 16 * 16
    if n16 <= 256:
        bn = n16
    else:
        bn = min((256, 192, 128), key=lambda b: ((n + b - 1) // b * b - n, -b))
    tiles_m, tiles_n, num_kb = (m + 127) // 128, (n + bn - 1) // bn, (k + 63) // 64
    return tiles_m >= 2 and tiles_m * tiles_n >= sms and bn % 32 == 0 and bn >= 128 and num_kb >= 3


def time_op_classes(plan, torch, detail=None):
    """Per-op CUDA-event timing of one plan (eager, after warm-up) -> {class: [ms, flops, bytes, launches]}."""
    import ctypes
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    res = {}
    evs = []
    for op in plan.ops:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = op.fn(*op.args, stream)
        e1.record()
        assert rc == 0
        evs.append((op, e0, e1))
    torch.cuda.synchronize()
    for op, e0, e1 in evs:
        ms = e0.elapsed_time(e1)
        fl = by = 0
        a = op.args
        if op.name == "uavsal_pw_gemm":
            m, k, n = a[3], a[4], a[7]
            fl = 2.0 * m * k * n
            by = 4.0 * m * k + (2.0 if a[9] & 32 else 4.0) * m * n + 4.0 * n * k          # UAVSAL_F_OUT_Q16: 16-bit hidden rows
            if _pw_is_pair(m, k, n) and not ((a[9] & 2) and n >= 64):      # the launcher's rule (gemm_tc.cu want_cluster), not EPI_RES
                # two instantiations: EPI_Q16 (the 256 -> 1536 class of expand convs, 16-bit hidden rows out) and EPI_STD
                r = res.setdefault("uavsal_pw_gemm/pair_q16" if a[9] & 32 else "uavsal_pw_gemm/pair", [0.0, 0.0, 0.0, 0])
                r[0] += ms; r[1] += fl; r[2] += by; r[3] += 1
        elif op.name == "uavsal_conv3x3":
            nimg, hh, ww, c, cout = a[3], a[4], a[5], a[6], a[8]
            fl = 2.0 * nimg * hh * ww * 9 * c * cout
            by = 4.0 * nimg * hh * ww * (c + cout)
        elif op.name == "uavsal_twa_sequence":
            t_steps, hh, ww, c, batch = a[6], a[7], a[8], a[9], a[17]
            fl = 2.0 * batch * t_steps * hh * ww * 9 * 2 * c * c
            by = 4.0 * batch * t_steps * hh * ww * 3 * c
        elif op.name == "uavsal_dw_project":
            nimg, hh, ww, hidden, cout = a[2], a[3], a[4], a[5], a[10]
            fl = 2.0 * nimg * hh * ww * hidden * cout + 18.0 * nimg * hh * ww * hidden
            by = nimg * hh * ww * ((2.0 if a[12] & 64 else 4.0) * hidden + 4.0 * cout)     # UAVSAL_F_HID_Q16
        elif op.name == "uavsal_expand_dw3x3":
            nimg, hh, ww, cin, hidden, stride = a[3], a[4], a[5], a[6], a[10], a[11]
            ho, wo = (hh if stride == 1 else (hh - 1) // 2 + 1), (ww if stride == 1 else (ww - 1) // 2 + 1)
            fl = 2.0 * nimg * hh * ww * cin * hidden + 18.0 * nimg * ho * wo * hidden
            by = 4.0 * nimg * (hh * ww * cin + ho * wo * hidden)
        elif op.name == "uavsal_dw3x3":
            nimg, hh, ww, c, stride = a[3], a[4], a[5], a[6], a[7]
            ho, wo = (hh if stride == 1 else (hh - 1) // 2 + 1), (ww if stride == 1 else (ww - 1) // 2 + 1)
            fl = 18.0 * nimg * ho * wo * c
            by = nimg * c * ((2.0 if a[1] == -2 else 4.0) * hh * ww + 4.0 * ho * wo)          # UAVSAL_PLANE_Q16 input
        r = res.setdefault(op.name, [0.0, 0.0, 0.0, 0])
        r[0] += ms; r[1] += fl; r[2] += by; r[3] += 1
        if detail is not None:
            detail.append("%-26s %-22s %8.1f us %8.1f TF/s %8.1f GB/s  %s" % (op.name, op.tag, ms * 1e3, fl / ms / 1e9 if ms else 0,
                                                                          by / ms / 1e6 if ms else 0, str(op.args[3:9])))
    return res


def bench_convlstm_config3(torch, dev, tens_peak, reps=3):
    """BASELINE config #3: ConvLSTM((45,80), 256 -> 256, 3x3, bias=False), batch 8, 64 time steps through
    uavsal_convlstm_sequence (one implicit-GEMM launch per step, cell update fused).  CUDA events around the sequence op."""
    import ctypes
    from iip_uavsal_saliency_b200.engine import Plan, W
    torch.manual_seed(0)
    b, t, c = 8, 64, 256
    plan = Plan(dev, 3, "tc")
    w4 = torch.empty(4 * c, 2 * c, 3, 3, device=dev)
    torch.nn.init.xavier_uniform_(w4)                               # model_convlstm.py:109
    x, h0, seq = plan.alloc(b * t * MH * MW, c), plan.alloc(b * MH * MW, c), plan.alloc(b * t * MH * MW, c)
    x.t.copy_(torch.randn(2, 64, c, device=dev).to(torch.bfloat16).repeat(1, b * t * MH * MW // 64, 1) * torch.tensor([1.0, 0.004], device=dev).view(2, 1, 1).to(torch.bfloat16))
    cst = plan.tensor((b, MH * MW, c))
    plan.lstm(x, h0, cst, b, t, MH, MW, c, c, W(w4), None, seq, tag="config3")
    op = plan.ops[-1]
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    ts = []
    for i in range(reps + 1):
        cst.zero_()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        assert op.fn(*op.args, stream) == 0
        e1.record()
        torch.cuda.synchronize()
        if i:
            ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * b * MH * MW * (4 * c) * (9 * 2 * c) * t
    tf = flops / ms / 1e9
    finite = bool(torch.isfinite(seq.t[:, -MH * MW:].float()).all().item())
    return {"workload": "ConvLSTM gate conv + cell update, hidden 256 ch, 45x80, batch 8, 64 time steps (BASELINE config #3), bf16x3 exact mode",
            "ms": round(ms, 3), "ms_per_step": round(ms / t, 4), "flops_algorithmic": flops, "tflops_alg": round(tf, 1),
            "frac": round(tf / tens_peak, 4), "issued_frac": round(3 * tf / tens_peak, 4), "peak_tflops": tens_peak,
            "kernel": "gemm_tc2_kernel<MODE_CONV,EPI_LSTM,TERMS=3,CL=2> x 64 launches", "reps": reps, "outputs_finite": finite,
            "inputs_bytes": int(x.t.numel() * 2 + seq.t.numel() * 2), "parity": "tests/test_gpu_parity.py::test_convlstm_config3_shape_pair_mode"}


def bench_metrics_config5(torch, dev, rank, world, hbm_peak, D, total_pairs=16384, reps=5):
    """BASELINE config #5: CC/NSS/KLD/SIM on 16 384 synthetic 360x640 pairs (fp32, as evalscores_vid_torch feeds them), pairs
    i -> rank i % world, one fused launch per rank, then the design's only collective: all_reduce(SUM) of [sum CC, sum NSS, sum KLD,
    sum SIM, n_valid] (NCCL when world > 1).  Pair i is base pair i % 64 of a seeded set, so the dataset means are known from 64 pairs."""
    from iip_uavsal_saliency_b200 import synth, utils_score_torch as US
    base_p, base_t = synth.make_metric_pairs(64, H, W, seed=0)
    mine = D.shard_indices(total_pairs, rank, world)
    bp, bt = torch.from_numpy(base_p).to(dev), torch.from_numpy(base_t).to(dev)
    sel = torch.tensor([i % 64 for i in mine], device=dev)
    pred, true = bp[sel].contiguous(), bt[sel].contiguous()           # resident: 2.76 MB per pair (45 GB at one rank) >> L2
    del bp, bt
    vals = US.metrics4(pred, true)
    torch.cuda.synchronize()
    ts, tk = [], []
    means = None
    for _ in range(reps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record()
        vals = US.metrics4(pred, true)
        e1.record()
        means = D.allreduce_metric_means(D.metric_partial(vals))
        e2.record()
        torch.cuda.synchronize()
        ts.append(D.max_over_ranks(e0.elapsed_time(e2) / 1e3, dev))
        tk.append(D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev))
    ts.sort(); tk.sort()
    t, tker = ts[len(ts) // 2], tk[len(tk) // 2]
    # check: the all-reduced dataset means equal the means of the 64 base pairs (every base pair occurs equally often)
    ref64 = US.metrics4(torch.from_numpy(base_p).to(dev), torch.from_numpy(base_t).to(dev)).double().mean(0)
    rel = ((means.double().to(dev) - ref64).abs() / ref64.abs()).max().item() if total_pairs % 64 == 0 else None
    by = 3.0 * H * W * 4
    gbs = by * len(mine) / tker / 1e9
    out = {"workload": "CC+NSS+KLD+SIM on %d synthetic 360x640 fp32 pairs (BASELINE config #5), pairs i -> rank i %% world, final all-reduce of the 5-vector" % total_pairs,
           "pairs": total_pairs, "pairs_per_rank": len(mine), "pairs_per_s": total_pairs / t, "ms": round(1e3 * t, 3), "kernel_ms": round(1e3 * tker, 3),
           "allreduce": ("nccl" if world > 1 else "none (one rank)"), "allreduce_ms": round(1e3 * (t - tker), 3),
           "bytes_per_pair": by, "gbs_per_gpu": round(gbs, 1), "peak_gbs": hbm_peak, "frac": round(gbs / hbm_peak, 4),
           "means": {k: float(v) for k, v in zip(("CC", "NSS", "KLD", "SIM"), means.tolist())},
           "means_vs_single_rank_rel_err": rel, "reps": reps,
           "inputs": "%d pairs per rank resident in HBM (%.1f GB, 64 distinct pairs tiled) - larger than L2" % (len(mine), by * len(mine) / 1e9)}
    if rel is not None:
        assert rel < 1e-5, "sharded metric means differ from the single-rank means: %g" % rel
    del pred, true
    torch.cuda.empty_cache()
    return out, (base_p, base_t)


def run_product_arm(args):
    import numpy as np
    import torch
    from iip_uavsal_saliency_b200 import _ext, dist as D
    from iip_uavsal_saliency_b200.model import UAVSal
    from iip_uavsal_saliency_b200.runner import ClipRunner
    from iip_uavsal_saliency_b200 import synth   # seeded synthetic inputs / weights

    t_start = time.perf_counter()
    # a freshly provisioned box was seen to fail one CUDA driver initialisation (a failed cuInit can stick to the process): probe
    # in a child process first, retrying, and touch CUDA here only once the probe has succeeded
    probe = "import sys, torch; sys.exit(0 if torch.cuda.is_available() else 1)"
    for attempt in range(4):
        if subprocess.run([sys.executable, "-c", probe], capture_output=True).returncode == 0:
            break
        time.sleep(5.0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py product arm needs a B200; the sm_100a library has no CPU fallback")
    rank, world, local = D.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _ext.check(_ext.load().uavsal_device_ok(local), "device_ok")
    tens_peak, hbm_peak, peak_src = load_peaks()

    model = UAVSal().eval()
    model.load_state_dict(synth.make_state_dict("lively", 0), strict=True)
    model = model.to(dev).set_mode(precision=args.precision)
    gauss, ob = load_priors()
    runner = ClipRunner(model, gauss, ob, batch_size=BATCH, out_hw=(H, W), use_graph=not args.no_graph, depth=args.depth,
                        clip_backbone=not args.per_call_backbone, single_stream=args.single_stream, whole_clip=not args.per_call,
                        clips_per_plan=args.clips_per_plan)
    sampler = ClockSampler(local)           # started ahead of the plan build: nvidia-smi is streaming by the time the timed region begins
    sampler.start()
    runner.warm(FRAMES, H, W)

    # this rank's share of a step's clips: clip c -> rank c % world.  The clips of a step are drawn in rotation from a pool of
    # distinct synthetic clips (44 MB each) larger than the 126 MB L2; the arena traffic of a clip (GBs) exceeds it by far anyway
    my_ids = [c for c in range(args.clips) if c % world == rank]
    n_distinct = max(4, min(len(my_ids), 8))                           # >= 4 x 44 MB: more than the 126 MB L2 at every N
    host_pool = [torch.from_numpy(synth.make_clip(100 + 16 * rank + i, FRAMES, H, W)).pin_memory() for i in range(n_distinct)]
    dev_pool = [c.to(dev) for c in host_pool]
    n_mine = len(my_ids)
    host_out = [torch.empty((OUT_PER_CLIP, H, W), dtype=torch.uint8).pin_memory() for _ in range(2)]
    dev_out = [torch.empty((OUT_PER_CLIP, H, W), dtype=torch.uint8, device=dev) for _ in range(2)]

    # the runner pipelines calls (and clips) over its own streams; finish() joins them into the timed stream
    def make_step(pool, outs, per_step):
        def step(i):
            for c in range(per_step):
                k = i * per_step + c
                runner.run_clip(pool[k % len(pool)], want_maps=False, out=outs[k % 2], sync=False)
        return step

    def timed(fn, steps, warmup, sampler=None):
        for i in range(warmup):
            fn(i)
        runner.finish()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if sampler is not None:
            sampler.mark_begin()
        e0.record()
        for i in range(steps):
            fn(warmup + i)
        runner.finish()
        e1.record()
        torch.cuda.synchronize()
        if sampler is not None:
            sampler.mark_end()
        if world > 1:
            torch.distributed.barrier()
        return D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev)

    warm = max(3, args.warmup)
    t_res = timed(make_step(dev_pool, dev_out, n_mine), args.steps, warm, sampler)
    clocks = sampler.stop()
    t_e2e = timed(make_step(host_pool, host_out, n_mine), args.steps, warm)
    frames_per_step = args.clips * OUT_PER_CLIP                       # all ranks together
    value = frames_per_step * args.steps / t_res
    e2e = frames_per_step * args.steps / t_e2e
    # secondary: weak scaling - a fixed number of clips per GPU per step (what round 1 reported as the headline)
    wk_clips, wk_steps = 2, max(3, min(args.steps, 10))
    t_weak = timed(make_step(dev_pool, dev_out, wk_clips), wk_steps, 2)
    weak = {"value": world * wk_clips * OUT_PER_CLIP * wk_steps / t_weak, "unit": "frames/s", "clips_per_gpu_per_step": wk_clips, "steps": wk_steps, "scaling": "weak"}

    cpp = args.clips_per_plan if not args.per_call else 1
    if not args.per_call:
        # launches of this rank in the timed region: full batches of cpp clips plus single-clip plans for a remainder
        nfull, nrem = divmod(n_mine * args.steps, cpp)
        l_full = runner._plan(OUT_PER_CLIP * cpp, H, W, 0, "all", BATCH * T, cpp).num_launches
        l_one = runner._plan(OUT_PER_CLIP, H, W, 0, "all", BATCH * T).num_launches
        my_launches = nfull * l_full + nrem * l_one
    else:
        calls = [20, 20, 20]
        per_clip = sum(runner._plan(n, H, W, 0).num_launches for n in calls) + (runner._plan(OUT_PER_CLIP, H, W, 0, "sfnet").num_launches if not args.per_call_backbone else 0)
        my_launches = per_clip * n_mine * args.steps
    launches = int(D.sum_over_ranks(float(my_launches), dev)) if world > 1 else int(my_launches)

    # ---- config #5 on every rank (sharded, NCCL all-reduce), config #3 + rooflines + CPU legs on rank 0 ----
    metrics5 = convlstm3 = None
    base_pairs = None
    if not args.skip_aux:
        metrics5, base_pairs = bench_metrics_config5(torch, dev, rank, world, hbm_peak, D, args.metric_pairs)
    if rank != 0:
        D.shutdown()
        return 0
    if not args.skip_aux:
        convlstm3 = bench_convlstm_config3(torch, dev, tens_peak)
        torch.cuda.empty_cache()

    # per-op CUDA-event timing of the plan the runner actually replays (two clips = 120 frames: six 20-frame reference calls)
    if args.per_call:
        plan_m = model.get_plan(dev, 20, H, W, x_kind=2, post_hw=(H, W), cb_shared=True)
        prof_frames = 20
    else:
        plan_m = runner._plan(OUT_PER_CLIP * cpp, H, W, 0, "all", BATCH * T, cpp)
        prof_frames = OUT_PER_CLIP * cpp
    for ci in range(prof_frames // min(prof_frames, OUT_PER_CLIP)):
        nfr = min(prof_frames, OUT_PER_CLIP)
        plan_m.named["x_in"][ci * nfr:(ci + 1) * nfr].copy_(dev_pool[ci % len(dev_pool)][:nfr])
    cls = time_op_classes(plan_m, torch)
    detail = []
    cls = time_op_classes(plan_m, torch, detail)
    if args.dump_ops:
        with open(args.dump_ops, "w") as fh:
            fh.write("\n".join(detail) + "\n")
    tot_ms = sum(v[0] for k, v in cls.items() if "/" not in k)                 # "name/sub" entries are subsets of "name"
    breakdown = {k: {"ms": round(v[0], 3), "share": round(v[0] / tot_ms, 3), "launches": v[3],
                     "tflops": round(v[1] / v[0] / 1e9, 1) if v[0] and v[1] else None,
                     "gbs": round(v[2] / v[0] / 1e6, 1) if v[0] and v[2] else None}
                 for k, v in sorted(cls.items(), key=lambda kv: -kv[1][0]) if "/" not in k}
    traffic = load_traffic()

    def per_launch_traffic(fam):
        t = traffic.get("kernels", {}).get(fam)
        return t.get("dram_bytes_per_launch") if t else None

    terms = 3 if args.precision == "exact" else 1
    # dominant kernel = the top entry of the ncu launch list: gemm_tc2_kernel<MODE_PW, EPI_Q16 | EPI_STD, TERMS, CL=2> (the cta_group::2
    # pointwise GEMM of the wide layers; the instantiation writing 16-bit hidden rows carries the 256 -> 1536 class of expand convs);
    # the whole pointwise-GEMM family, small HBM-bound layers included, is reported next to it
    pairs = [k for k in ("uavsal_pw_gemm/pair_q16", "uavsal_pw_gemm/pair") if cls.get(k, [0])[0] > 0]
    dom = max(pairs, key=lambda k: cls[k][0]) if pairs else "uavsal_pw_gemm"
    epi = "EPI_Q16" if dom.endswith("q16") else "EPI_STD"
    ach = cls[dom][1] / cls[dom][0] / 1e9
    roofline = {"kernel": "gemm_tc2_kernel<MODE_PW,%s,TERMS=%d,CL=2> (cta_group::2 tcgen05 pointwise-conv GEMM, %d launches per %d-frame plan)" % (epi, terms, cls[dom][3], prof_frames),
                "bound": "tensor", "achieved": round(ach, 2), "peak": tens_peak, "unit": "TFLOP/s", "frac": round(ach / tens_peak, 4),
                "traffic": per_launch_traffic(dom), "traffic_source": traffic.get("source"), "peak_source": peak_src,
                "algorithmic_flops_per_launch": round(cls[dom][1] / cls[dom][3]), "avg_launch_us": round(1e3 * cls[dom][0] / cls[dom][3], 2),
                "issued_frac": round(terms * ach / tens_peak, 4),
                "note": "achieved = algorithmic 2*M*K*N flops (1x) summed over the kernel's launches / summed CUDA-event time, measured in this run; the "
                        "bf16x3 split issues 3x that on the tensor pipe (issued_frac); traffic = ncu dram bytes per launch from this round's capture"}
    allpw = "uavsal_pw_gemm"
    ach_a = cls[allpw][1] / cls[allpw][0] / 1e9
    roofline_all_pw = {"kernel": "every pointwise-conv GEMM launch (CL=1|2, EPI_STD|EPI_RES|EPI_Q16; %d launches, the small-K backbone layers are HBM-bound)" % cls[allpw][3],
                       "bound": "tensor", "achieved": round(ach_a, 2), "peak": tens_peak, "unit": "TFLOP/s", "frac": round(ach_a / tens_peak, 4),
                       "issued_frac": round(terms * ach_a / tens_peak, 4), "gbs": round(cls[allpw][2] / cls[allpw][0] / 1e6, 1)}
    hb = "uavsal_dw3x3"
    ach_h = cls[hb][2] / cls[hb][0] / 1e6
    roofline_hbm = {"kernel": "dw3x3_tma_kernel / dw3x3_kernel (depthwise 3x3 + BN + ReLU6, %d launches per %d-frame plan)" % (cls[hb][3], prof_frames),
                    "bound": "hbm", "achieved": round(ach_h, 1), "peak": hbm_peak, "unit": "GB/s", "frac": round(ach_h / hbm_peak, 4),
                    "traffic": per_launch_traffic(hb), "algorithmic_bytes_per_launch": round(cls[hb][2] / cls[hb][3]),
                    "avg_launch_us": round(1e3 * cls[hb][0] / cls[hb][3], 2)}
    whole = {"tflops_alg": round(54.13e9 * prof_frames / (tot_ms * 1e-3) / 1e12, 1), "frac_of_tensor_peak": round(54.13e9 * prof_frames / (tot_ms * 1e-3) / 1e12 / tens_peak, 4),
             "sum_of_kernel_ms_per_plan": round(tot_ms, 3), "flops_per_frame": 54.13e9}

    # ---- CPU baseline on this box's host cores (N = 1 only), bounded samples of the same workloads ----
    cpu_baseline = None
    if world == 1 and not args.skip_cpu:
        arm = CpuArm()
        clip_np = host_pool[0].numpy()
        arm.uavsal_call(clip_np, 5)
        t_cpu = arm.uavsal_call(clip_np, 20)
        cpu_baseline = {"value": 20 / t_cpu, "unit": "frames/s", "cores": arm.cores, "kind": arm.kind,
                        "sample": "one 20-frame call (batch_size=4 x time_dims=5) incl. normalisation and CPU post-process, %.1f s" % t_cpu}
        if convlstm3 is not None:
            s3 = arm.convlstm_steps(2)
            convlstm3["cpu_baseline"] = {"ms_per_step": round(1e3 * s3, 1), "ms_64_steps_extrapolated": round(64e3 * s3, 1), "cores": arm.cores, "kind": arm.kind,
                                         "sample": "2 steps at batch 8 timed, x 32"}
        if metrics5 is not None:
            t5 = arm.metrics_pairs(base_pairs[0][:32], base_pairs[1][:32])
            metrics5["cpu_baseline"] = {"pairs_per_s": round(32 / t5, 1), "cores": arm.cores, "kind": arm.kind, "sample": "32 pairs, the four metric functions called one after the other"}
    if metrics5 is not None:
        tr = traffic.get("kernels", {}).get("metrics4")
        metrics5["dram_traffic_ratio"] = tr.get("dram_read_ratio") if tr else None
        metrics5["traffic_source"] = tr.get("source") if tr else None

    line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * t_res / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "bf16x3-split (fp32 accumulate)" if args.precision == "exact" else "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "clips_per_step": args.clips, "partition": "clip c -> rank c % world", "frames_in_per_clip": FRAMES,
                       "maps_out_per_clip": OUT_PER_CLIP, "precision": args.precision, "cuda_graph": not args.no_graph, "calls_in_flight": args.depth,
                       "clips_per_plan": args.clips_per_plan,
                       "hidden_rows": "16-bit fixed point of the ReLU6 output (|err| <= 4.6e-5) for hidden tensors >= 1152 channels, fp32 otherwise"
                                      if os.environ.get("UAVSAL_HIDDEN_Q16", "1") != "0" else "fp32",
                       "plan": "one per %d clip(s) (60 frames each, call size 20 passed to the call-granular kernels)" % cpp if not args.per_call else "one per 20-frame call",
                       "l2": "inputs larger than L2: %d distinct 44 MB clips per rank rotated, and ~4 GB of arena traffic per clip" % n_distinct},
            "timed_region_s": round(t_res, 3),
            "e2e": {"value": e2e, "unit": "frames/s", "h2d_bytes_per_step": args.clips * OUT_PER_CLIP * H * W * 3,
                    "d2h_bytes_per_step": args.clips * OUT_PER_CLIP * H * W, "timed_region_s": round(t_e2e, 3)},
            "weak_scaling": weak,
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "roofline_all_pw": roofline_all_pw, "roofline_hbm": roofline_hbm,
            "whole_network": whole, "cpu_baseline": cpu_baseline, "convlstm_config3": convlstm3, "metrics_config5": metrics5,
            "breakdown_per_plan": breakdown, "breakdown_frames": prof_frames, "hbm_peak_gbs": hbm_peak,
            "arena_bytes_per_plan": int(plan_m.arena_bytes), "wall_s": round(time.perf_counter() - t_start, 1)}
    print(json.dumps(line), flush=True)
    D.shutdown()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--clips", type=int, default=16, help="clips per step over ALL ranks (clip c -> rank c %% world)")
    ap.add_argument("--precision", default="exact", choices=["exact", "fast"])
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--clips-per-plan", type=int, default=2, help="clips the runner queues into one plan (ConvTWA advances them as a batch); "
                    "1 = every clip on its own")
    ap.add_argument("--per-call", action="store_true", help="keep Demo_Test's loop of 20-frame calls instead of one plan per clip")
    ap.add_argument("--single-stream", action="store_true", help="queue all stages of all calls on one stream (no overlap)")
    ap.add_argument("--per-call-backbone", action="store_true", help="run the SRF-Net per 20-frame call (as Demo_Test does) instead of once per clip")
    ap.add_argument("--depth", type=int, default=2, help="calls in flight per GPU (ClipRunner stream pipelining; 1 = serial)")
    ap.add_argument("--dump-ops", default="", help="write per-op CUDA-event timings of one plan to this file")
    ap.add_argument("--skip-aux", action="store_true", help="skip the config #3 / #5 measurements")
    ap.add_argument("--skip-cpu", action="store_true", help="skip the CPU baseline legs")
    ap.add_argument("--metric-pairs", type=int, default=16384, help="config #5: total pairs over all ranks")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_product_arm(args)


if __name__ == "__main__":
    sys.exit(main())
