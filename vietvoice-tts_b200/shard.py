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

from .arch import FULL, ArchConfig


def chunk_cost(total_frames: int, arch: ArchConfig = FULL) -> float:
    """Algorithmic FLOPs of one DiT forward over `total_frames` tokens: depth x (linear layers + attention), read from
    the architecture (per token per layer 6 d^2 (QKV) + 2 d^2 (out) + 4 d ff (FFN) + 4 T d (attention))."""
    t = float(total_frames)
    d, ff = arch.dim, arch.ff_dim
    lin = arch.depth * (8 * d * d + 4 * d * ff)
    att = arch.depth * 4 * d
    return t * (lin + att * t)


def assign_chunks(total_frames: Sequence[int], world: int, arch: ArchConfig = FULL) -> List[List[int]]:
    """-> per-rank list of chunk indices (ascending).  Deterministic: ties broken by index, then by rank."""
    order = sorted(range(len(total_frames)), key=lambda i: (-chunk_cost(total_frames[i], arch), i))
    load = [0.0] * world
    out: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: (load[k], k))
        out[r].append(i)
        load[r] += chunk_cost(total_frames[i], arch)
    return [sorted(x) for x in out]


def imbalance(total_frames: Sequence[int], world: int, arch: ArchConfig = FULL) -> float:
    """max over ranks of the assigned cost / mean cost (1.0 = perfectly balanced): the LPT part of a strong-scaling
    loss, before any GPU effect."""
    parts = assign_chunks(total_frames, world, arch)
    loads = [sum(chunk_cost(total_frames[i], arch) for i in p) for p in parts]
    mean = sum(loads) / max(world, 1)
    return max(loads) / mean if mean > 0 else 1.0


def dispatch_requests(costs: Sequence[float], world: int) -> List[int]:
    """Owner rank of each request of a stream, decided in ARRIVAL order (an online front-end balancer: request i goes to
    the rank with the least estimated work accepted so far; ties to the lowest rank).  Deterministic, so every rank of a
    job that replays the same stream computes the same owners without talking to the others.  Against `i % world` it
    keeps the slowest rank's makespan near the mean when request sizes and NFE differ by an order of magnitude."""
    load = [0.0] * max(world, 1)
    owners: List[int] = []
    for c in costs:
        r = min(range(len(load)), key=lambda k: (load[k], k))
        owners.append(r)
        load[r] += float(c)
    return owners


class Sharder:
    def __init__(self, rank: int = 0, world: int = 1, group=None, arch: ArchConfig = FULL):
        self.rank, self.world, self.group, self.arch = rank, world, group, arch

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
        return assign_chunks(total_frames, self.world, self.arch)[self.rank]

    def gather(self, local: Dict[int, np.ndarray], n_chunks: int, error: Optional[str] = None) -> Dict[int, np.ndarray]:
        """All ranks receive every chunk's waveform, keyed by chunk index.  A rank whose synthesis failed still takes
        part, passing `error`: the failure is then raised on EVERY rank instead of leaving the others blocked in the
        collective until its timeout."""
        if self.world == 1:
            if error is not None:
                raise RuntimeError(error)
            return dict(local)
        import torch
        import torch.distributed as dist
        mine: dict = {int(k): np.asarray(v) for k, v in local.items()}
        # Phase 1 — one small tensor per rank: sample count and rank (ndim) of every chunk it holds, plus a status word
        # (0 = int16 waves with unit leading dimensions: the fast path, 1 = something else: pickle it, 2 = failed).
        meta = torch.full((2 * n_chunks + 1,), -1, dtype=torch.int64)
        plain = all(0 <= k < n_chunks and v.dtype == np.int16 and v.ndim >= 1 and v.size == v.shape[-1]
                    for k, v in mine.items())
        for k, v in mine.items():
            if 0 <= k < n_chunks:
                meta[k], meta[n_chunks + k] = v.shape[-1], v.ndim
        meta[2 * n_chunks] = 2 if error is not None else (0 if plain else 1)
        metas = [torch.empty_like(meta) for _ in range(self.world)]
        dist.all_gather(metas, meta, group=self.group)
        status = max(int(m[2 * n_chunks]) for m in metas)
        merged: Dict[int, np.ndarray] = {}
        if status == 0:
            # Phase 2 — the waves themselves as ONE int16 tensor per rank (chunks in ascending index order, padded to the
            # largest rank): no pickling, one collective (at 8 ranks / 349 s of audio the pickled form took ~20 ms).
            sizes = [int(m[:n_chunks].clamp(min=0).sum()) for m in metas]
            width = max(max(sizes), 1)
            buf = torch.zeros(width, dtype=torch.int16)
            off = 0
            for k in sorted(mine):
                n = mine[k].shape[-1]
                buf[off:off + n] = torch.from_numpy(np.ascontiguousarray(mine[k]).reshape(-1))
                off += n
            outs = [torch.empty(width, dtype=torch.int16) for _ in range(self.world)]
            # (gloo has no int16: the same memory travels as bytes)
            dist.all_gather([o.view(torch.uint8) for o in outs], buf.view(torch.uint8), group=self.group)
            for m, o in zip(metas, outs):
                flat, off = o.numpy(), 0
                for k in range(n_chunks):
                    n = int(m[k])
                    if n >= 0:
                        merged[k] = flat[off:off + n].reshape((1,) * (int(m[n_chunks + k]) - 1) + (n,)).copy()
                        off += n
        else:
            parts: List[Optional[dict]] = [None] * self.world
            if error is not None:
                mine = {"__error__": f"rank {self.rank}: {error}"}
            dist.all_gather_object(parts, mine, group=self.group)
            failed = [p["__error__"] for p in parts if p and "__error__" in p]
            if failed:
                raise RuntimeError("; ".join(failed))
            for p in parts:
                merged.update(p or {})
        missing = [i for i in range(n_chunks) if i not in merged]
        if missing:
            raise RuntimeError(f"chunks {missing} were not produced by any rank")
        return merged
