"""Multi-GPU plumbing: envs shard trivially (no per-step communication), so the only
collective is an all-gather of the per-shard episode statistics vector
(SURVEY.md section 8(e)).  One process per GPU, ``torch.distributed`` (NCCL on GPUs,
gloo in the CPU tests)."""
from typing import Dict, List, Sequence

import numpy as np

STATS_KEYS = ("episodes", "steps", "sum_makespan", "min_makespan", "max_makespan", "sum_return",
              "envs_done", "envs_error")
_I64_MAX = np.iinfo(np.int64).max


def shard_range(total_envs: int, rank: int, world_size: int):
    """Contiguous shard [lo, hi) of the global env ids owned by `rank` (remainder to low ranks)."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def combine_stats(per_rank: Sequence[Sequence[int]]) -> Dict[str, int]:
    a = np.asarray(per_rank, dtype=np.int64).reshape(-1, len(STATS_KEYS))
    out = {k: int(a[:, i].sum()) for i, k in enumerate(STATS_KEYS)}
    out["min_makespan"] = int(a[:, 3].min())
    out["max_makespan"] = int(a[:, 4].max())
    if out["min_makespan"] == _I64_MAX:
        out["min_makespan"] = -1   # no finished episode anywhere
    return out


def all_gather_stats(local_stats: Dict[str, int], device=None) -> Dict[str, int]:
    """All-gather the 8 x int64 stats vector of every rank and reduce it on each rank."""
    import torch
    import torch.distributed as dist
    vec = [int(local_stats[k]) for k in STATS_KEYS]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return combine_stats([vec])
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else "cpu"
    t = torch.tensor(vec, dtype=torch.int64, device=device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return combine_stats([o.cpu().tolist() for o in out])
