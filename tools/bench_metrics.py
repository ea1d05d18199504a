"""BASELINE config #5: utils_score_torch CC/NSS/KLD/SIM on synthetic 360x640 saliency / fixation map pairs, sharded by pair
over the ranks with one final all-reduce (dist.allreduce_metric_means).

    python tools/bench_metrics.py [--pairs 2048] [--dtype f32|u8]            # one GPU
    torchrun --nproc-per-node N ... tools/bench_metrics.py --pairs 16384      # pairs r::N per rank

The pairs are tiled from 64 distinct synthetic pairs generated on the host (oracle.synth.make_metric_pairs - the seeded
generator, not the checker).  Prints one JSON line on rank 0: pairs/s, achieved GB/s over the algorithmic bytes
(3 planes x H x W x sizeof(dtype), read once) against the measured HBM peak, and the dataset means."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from iip_uavsal_saliency_b200 import dist as D
from iip_uavsal_saliency_b200 import utils_score_torch as US
from oracle import synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2048, help="total pairs over all ranks")
    ap.add_argument("--dtype", default="f32", choices=["f32", "u8"])
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = D.init_process_group()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    H, W = 360, 640
    mine = len(D.shard_indices(args.pairs, rank, world))
    base_p, base_t = synth.make_metric_pairs(64, H, W, seed=rank)
    bp, bt = torch.from_numpy(base_p).to(dev), torch.from_numpy(base_t).to(dev)
    if args.dtype == "u8":
        bp, bt = bp.round().clamp(0, 255).to(torch.uint8), bt.round().clamp(0, 255).to(torch.uint8)
    reps = (mine + 63) // 64
    pred = bp.repeat(reps, 1, 1, 1)[:mine].contiguous()
    true = bt.repeat(reps, 1, 1, 1)[:mine].contiguous()
    esz = pred.element_size()
    vals = US.metrics4(pred, true)
    torch.cuda.synchronize()
    ts = []
    for _ in range(args.reps):
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        vals = US.metrics4(pred, true)
        means = D.allreduce_metric_means(D.metric_partial(vals))
        e1.record()
        torch.cuda.synchronize()
        ts.append(D.max_over_ranks(e0.elapsed_time(e1) / 1e3, dev))
    ts.sort()
    t = ts[len(ts) // 2]
    if rank == 0:
        peak = 6551.4
        try:
            peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        by = 3.0 * H * W * esz
        gbs = by * mine / t / 1e9                    # per GPU (every rank holds the same number of pairs +-1)
        print(json.dumps({"metric": "saliency metric pairs/s (CC+NSS+KLD+SIM, 360x640)", "value": args.pairs / t, "unit": "pairs/s", "n_gpus": world,
                          "pairs": args.pairs, "dtype": args.dtype, "ms": 1e3 * t, "bytes_per_pair": by,
                          "roofline": {"bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4)},
                          "means": {k: float(v) for k, v in zip(("CC", "NSS", "KLD", "SIM"), means.tolist())},
                          "inputs": "%d pairs per rank (64 distinct, tiled) resident in HBM: %.1f GB > L2" % (mine, by * mine / 1e9)}), flush=True)
    D.shutdown()


if __name__ == "__main__":
    main()
