"""Multi-GPU plumbing for the denoise loop: one process per GPU, panels sharded contiguously, NO collective on
the per-step path (every panel — and its CFG pair — is an independent sample through the UNet).
torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in the CPU tests) is used only for the final
all-gather of latents and for the barrier / max-over-ranks timing in bench.py.
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced split of ``total`` panels: the first ``total % world`` ranks get one extra."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Var-res buckets (cfg3): greedy longest-processing-time assignment of panel indices to ranks so that
    sum(cost) per rank is balanced; cost ~ H*W (+ an N^2 self-attention term) per panel."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: loads[k])
        out[r].append(i)
        loads[r] += costs[i]
    return [sorted(x) for x in out]


def panel_cost(height: int, width: int) -> float:
    """Relative cost of one panel-step: linear terms ~ pixels, self-attention ~ tokens^2 at latent/2."""
    px = height * width
    n1 = px / 256.0
    return px / (1024.0 * 1024.0) * 0.89 + (n1 / 4096.0) ** 2 * 0.11        # cfg2 split: 11 % self-SDPA


def init_from_env(backend: str = "nccl") -> Tuple[int, int, int]:
    """(rank, world, local_rank) from the torchrun environment; initialises the default process group if
    WORLD_SIZE > 1.  MASTER_ADDR defaults to 127.0.0.1 (container hostnames may not resolve)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def gather_latents(local: torch.Tensor, counts: Sequence[int]) -> torch.Tensor:
    """All-gather the per-rank final latents (n_r, 4, h, w) into panel order on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    m = max(counts)
    pad = torch.zeros((m,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    return torch.cat([b[:c] for b, c in zip(bufs, counts)], dim=0)


def max_over_ranks(value: float, device) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
