"""Multi-GPU plumbing: one process per GPU, work sharded by independent clip / metric pair, no data-path
collective.  The only collective of the design is the final all-reduce of five scalars per rank
([sum CC, sum NSS, sum KLD, sum SIM, n_valid]) that turns per-rank metric sums into dataset means — the
multi-GPU counterpart of the per-video np.mean aggregation in utils_score_torch.py:563-581.

Within a clip the ConvTWA state, the temporal differences and the context prior couple frames, so a clip
never spans ranks."""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def init_process_group(backend: Optional[str] = None):
    """Initialise torch.distributed from the torchrun environment (no-op for a single process)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, rank=rank, world_size=world, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def shutdown():
    """Tear the process group down (quietens NCCL's leak warning at interpreter exit)."""
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


def shard_indices(n_items: int, rank: int, world: int) -> List[int]:
    """Static round-robin partition: item i -> rank i % world (clips of equal length)."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, n_items, world))


def metric_partial(values: torch.Tensor) -> torch.Tensor:
    """(n,4) per-pair metrics -> float64 [sum CC, sum NSS, sum KLD, sum SIM, n_valid]; rows containing NaN
    (the reference skips NaN scores with np.nanmean-style aggregation) are not counted."""
    v = values.double()
    ok = ~torch.isnan(v).any(1)
    out = torch.zeros(5, dtype=torch.float64, device=values.device)
    out[:4] = v[ok].sum(0)
    out[4] = ok.sum()
    return out


def allreduce_metric_means(partial: torch.Tensor) -> torch.Tensor:
    """All-reduce(SUM) the 5-vector over all ranks and return the four dataset means (float64, on every rank)."""
    total = partial.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM)
    return total[:4] / total[4].clamp(min=1.0)


def max_over_ranks(seconds: float, device=None) -> float:
    """Timing convention of bench.py: the job's time is the slowest rank's device time."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return seconds
    t = torch.tensor([seconds], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
