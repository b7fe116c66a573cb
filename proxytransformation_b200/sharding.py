"""Scene sharding of the preshape path over the GPUs of one box.

Scenes are independent units (eval-mode BatchNorm uses running statistics; the only batch coupling in the reference is
``torch.cat`` of equal-N scenes, preshape_norm_reverse_drop.py:427), so a batch is split into contiguous shards, one per
rank, with NO collective on the data path.  The single collective is an ``all_gather`` of a small fixed-size metric
tensor per rank at the end — the analogue of mmengine's ``collect_results`` behind ``GroundingMetric``
(embodiedscan/eval/metrics/grounding_metric.py:52-71).  Works with the ``nccl`` (CUDA tensors) and ``gloo`` (CPU
tensors, used by the tests) backends.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

METRIC_FIELDS = ("n_scenes", "survivors", "coord_checksum", "elapsed_ms", "launches")


def shard_range(n_scenes: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous ceil(n/world) scenes per rank: [start, stop).  Trailing ranks may be empty."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError(f"bad rank/world {rank}/{world}")
    per = -(-n_scenes // world)
    start = min(n_scenes, rank * per)
    return start, min(n_scenes, start + per)


def shard(items: Sequence, rank: int, world: int) -> Sequence:
    a, b = shard_range(len(items), rank, world)
    return items[a:b]


def scene_metrics(outputs: Sequence[torch.Tensor], elapsed_ms: float = 0.0, launches: int = 0, device=None) -> torch.Tensor:
    """Per-rank metric tensor [n_scenes, sum N', fp64 sum of all output coordinates, elapsed ms, kernel launches]."""
    dev = device if device is not None else (outputs[0].device if len(outputs) else "cpu")
    surv = float(sum(int(o.shape[0]) for o in outputs))
    chk = float(sum(o.double().sum().item() for o in outputs))
    return torch.tensor([float(len(outputs)), surv, chk, float(elapsed_ms), float(launches)], dtype=torch.float64, device=dev)


def gather_metrics(local: torch.Tensor) -> torch.Tensor:
    """all_gather of the per-rank metric tensors -> (world, len(METRIC_FIELDS)) on every rank (single-process: (1, F))."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local[None].clone()
    parts: List[torch.Tensor] = [torch.empty_like(local) for _ in range(dist.get_world_size())]
    dist.all_gather(parts, local)
    return torch.stack(parts)


def reduce_job(metrics: torch.Tensor) -> dict:
    """Whole-job numbers from the gathered table: scenes and survivors add up, time is the max over ranks."""
    m = metrics.cpu()
    return {"n_scenes": int(m[:, 0].sum().item()), "survivors": int(m[:, 1].sum().item()), "coord_checksum": m[:, 2].sum().item(),
            "elapsed_ms": m[:, 3].max().item(), "launches": int(m[:, 4].sum().item())}
