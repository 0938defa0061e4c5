"""Multi-GPU partitioning of independent chunks: one process per GPU, no data-path collective.

The unit of work is one (prompt, text-chunk) pair; chunks of a long text and separate requests are independent
(/root/reference/vietvoicetts/core/tts_engine.py:225-238 runs them sequentially with the same prompt).  Every rank
computes the SAME deterministic assignment (greedy, longest first, onto the least-loaded rank), synthesizes its own
chunks on its own GPU with its own weight replica, and the int16 waveforms are gathered on the host (gloo) in chunk
order for the reference's order-dependent cross-fade.  No NCCL / NVLink traffic on the hot path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np

# algorithmic FLOPs per token of one DiT forward: 22 x (16 d^2 + 4 T d) at d = 1024
_LIN = 22 * 16 * 1024 * 1024
_ATT = 22 * 4 * 1024


def chunk_cost(total_frames: int) -> float:
    t = float(total_frames)
    return t * (_LIN + _ATT * t)


def assign_chunks(total_frames: Sequence[int], world: int) -> List[List[int]]:
    """-> per-rank list of chunk indices (ascending).  Deterministic: ties broken by index, then by rank."""
    order = sorted(range(len(total_frames)), key=lambda i: (-chunk_cost(total_frames[i]), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += chunk_cost(total_frames[i])
    return [sorted(x) for x in out]


class Sharder:
    def __init__(self, rank: int = 0, world: int = 1, group=None):
        self.rank, self.world, self.group = rank, world, group

    @classmethod
    def from_torch_distributed(cls) -> "Sharder":
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return cls(0, 1, None)
        group = None
        if dist.get_backend() != "gloo":          # waveforms are host arrays: gather them over a gloo group
            group = dist.new_group(backend="gloo")
        return cls(dist.get_rank(), dist.get_world_size(), group)

    def assign(self, total_frames: Sequence[int]) -> List[int]:
        return assign_chunks(total_frames, self.world)[self.rank]

    def gather(self, local: Dict[int, np.ndarray], n_chunks: int) -> Dict[int, np.ndarray]:
        """All ranks receive every chunk's waveform, keyed by chunk index."""
        if self.world == 1:
            return dict(local)
        import torch.distributed as dist
        parts: List[Optional[dict]] = [None] * self.world
        dist.all_gather_object(parts, {int(k): np.asarray(v) for k, v in local.items()}, group=self.group)
        merged: Dict[int, np.ndarray] = {}
        for p in parts:
            merged.update(p or {})
        missing = [i for i in range(n_chunks) if i not in merged]
        if missing:
            raise RuntimeError(f"chunks {missing} were not produced by any rank")
        return merged
