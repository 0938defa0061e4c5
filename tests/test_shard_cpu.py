"""Multi-GPU host logic on CPU: deterministic chunk->rank assignment and the world_size-2 gloo gather of waveforms
(the only cross-rank step of the path)."""
import os
import socket

import numpy as np
import pytest

torch = pytest.importorskip("torch")
import torch.distributed as dist
import torch.multiprocessing as mp

from vietvoice_tts_b200.shard import Sharder, assign_chunks, chunk_cost


def test_assign_is_a_partition_and_balanced():
    rng = np.random.default_rng(0)
    frames = [int(x) for x in rng.integers(845, 1783, size=40)]          # SURVEY 8d cfg 3
    for world in (1, 2, 4, 8):
        parts = assign_chunks(frames, world)
        assert sorted(i for p in parts for i in p) == list(range(40))
        loads = [sum(chunk_cost(frames[i]) for i in p) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.08                  # greedy LPT stays within a few %
        assert parts == assign_chunks(frames, world)                      # deterministic
    assert assign_chunks([], 4) == [[], [], [], []]
    # the cost model is read from the architecture, not hard-coded for dim 1024 / 22 layers
    from vietvoice_tts_b200.arch import FULL, TINY
    assert chunk_cost(1000) == chunk_cost(1000, FULL) == 1000 * (22 * 16 * 1024 ** 2 + 22 * 4 * 1024 * 1000)
    assert chunk_cost(1000, TINY) == 1000 * (TINY.depth * (8 * TINY.dim ** 2 + 4 * TINY.dim * TINY.ff_dim)
                                             + TINY.depth * 4 * TINY.dim * 1000)
    assert assign_chunks([1000], 2) == [[0], []]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sh = Sharder.from_torch_distributed()
        frames = [900, 1700, 1200, 1500, 1000, 1782, 845]
        mine = sh.assign(frames)
        # stand-in for the per-rank GPU synthesis: a waveform that encodes (chunk index, length)
        local = {i: np.full((1, 1, (frames[i] - 563 - 1) * 256), i, dtype=np.int16) for i in mine}
        allw = sh.gather(local, len(frames))
        ok = sorted(allw) == list(range(len(frames))) and all(
            allw[i].dtype == np.int16 and allw[i].shape[-1] == (frames[i] - 564) * 256 and int(allw[i][0, 0, 0]) == i
            for i in range(len(frames)))
        # anything that is not a flat int16 wave (the engine never produces such, a caller might) takes the pickled path
        allf = sh.gather({i: np.full((2, 3), i, dtype=np.float32) for i in mine}, len(frames))
        ok = ok and sorted(allf) == list(range(len(frames))) and all(
            allf[i].dtype == np.float32 and allf[i].shape == (2, 3) and float(allf[i][1, 2]) == i for i in allf)
        # a rank whose synthesis failed still joins the collective: the error surfaces on EVERY rank at once
        # instead of leaving the healthy ranks blocked in the collective until the gloo timeout
        try:
            sh.gather({} if rank == 1 else local, len(frames), error="out of memory" if rank == 1 else None)
            ok = False
        except RuntimeError as exc:
            ok = ok and "rank 1: out of memory" in str(exc)
        q.put((rank, mine, ok))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_gather():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert all(ok for _, _, ok in res)
    assert sorted(res[0][1] + res[1][1]) == list(range(7)) and set(res[0][1]).isdisjoint(res[1][1])
