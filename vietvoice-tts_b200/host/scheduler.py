"""Request scheduler: concurrent synthesis requests -> micro-batches of chunks on one GPU (SURVEY.md 8f rank 3,
BASELINE config 5: a request stream sweeping voices and NFE 16 / 32 / 64).

The reference serves one request at a time: `TTSApi.synthesize_to_bytes` runs on an anyio worker thread, mutates the
shared `ModelConfig` (speed, seed) around the call and restores it afterwards
(/root/reference/vietvoicetts/api/tts_engine.py:64-69,79-91) — two concurrent requests race on that config, and each
request walks its chunks sequentially (core/tts_engine.py:225-238).  Here

  * `submit()` is thread-safe and returns a `concurrent.futures.Future`; `speed`, `nfe` and `seed` are per-request
    values that never touch the shared config;
  * one worker thread turns queued requests into chunks (`TTSEngine._prepare_inputs`, unchanged host logic), groups
    chunks by NFE (the sampling loop is one CUDA graph per (batch shape, nfe)), packs them longest-first into
    micro-batches of at most `max_batch_chunks` chunks / `max_batch_frames` mel frames and runs each micro-batch as ONE
    `Engine.synthesize_batch` call; chunks of different requests share a batch (rows are packed, no padding);
  * y0 of chunk c of a request is Philox(seed, c), so a request's waveform depends only on (text, voice, speed, nfe,
    seed) — not on what else was in flight — and equals what `TTSEngine.synthesize` returns for it;
  * with several ranks (one process per GPU) request i belongs to rank i % world: no data-path collective.

`plan_batches` is the pure packing function (unit-tested on the CPU).
"""
from __future__ import annotations

import queue
import threading
import time
from concurrent.futures import Future
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..shard import chunk_cost


@dataclass
class _Chunk:
    req: "_Request"
    index: int
    audio: np.ndarray
    ids: np.ndarray
    frames: int


@dataclass
class _Request:
    rid: int
    text: str
    voice: dict
    nfe: int
    speed: Optional[float]
    seed: int
    future: Future
    t0: float = field(default_factory=time.time)
    n_chunks: int = 0
    waves: Dict[int, np.ndarray] = field(default_factory=dict)


MAX_NFE = 100      # upper bound of ModelConfig.nfe_step (/root/reference/vietvoicetts/core/model_config.py:59-62)


def _resolve(fut: Future, result=None, error: Optional[BaseException] = None) -> None:
    """set_result / set_exception that tolerates a future the client cancelled or that is already resolved (an
    InvalidStateError here used to kill the worker thread and leave every later request hanging)"""
    try:
        if error is not None:
            fut.set_exception(error)
        else:
            fut.set_result(result)
    except Exception:
        pass


def plan_batches(frames: Sequence[int], max_chunks: int, max_frames: int) -> List[List[int]]:
    """Greedy longest-first packing of chunk indices into micro-batches: a batch closes when it holds `max_chunks`
    chunks or the next chunk would push it over `max_frames` mel frames.  Every index appears exactly once; a chunk
    longer than `max_frames` gets a batch of its own.  Sorting by length keeps chunks of similar cost together."""
    order = sorted(range(len(frames)), key=lambda i: (-chunk_cost(frames[i]), i))
    batches: List[List[int]] = []
    cur: List[int] = []
    cur_frames = 0
    for i in order:
        if cur and (len(cur) >= max_chunks or cur_frames + frames[i] > max_frames):
            batches.append(cur)
            cur, cur_frames = [], 0
        cur.append(i)
        cur_frames += frames[i]
    if cur:
        batches.append(cur)
    return batches


class RequestScheduler:
    def __init__(self, tts, max_batch_chunks: int = 8, max_batch_frames: int = 8 * 1800, max_wait_s: float = 0.004,
                 rank: int = 0, world: int = 1):
        """tts: a `vietvoice_tts_b200.host.tts_engine.TTSEngine` (its engine, text/audio processors and voice table
        are used; its config is read, never written)."""
        self.tts = tts
        self.max_batch_chunks = max_batch_chunks
        self.max_batch_frames = max_batch_frames
        self.max_wait_s = max_wait_s
        self.rank, self.world = rank, world
        self._q: "queue.Queue[Optional[_Request]]" = queue.Queue()
        self._closed = False
        self._next_id = 0
        self._id_lock = threading.Lock()
        self.batches_run = 0
        self.chunks_run = 0
        self._worker = threading.Thread(target=self._loop, name="vvb200-scheduler", daemon=True)
        self._worker.start()

    # ------------------------------------------------------------------------------------------ client side
    def owns(self, rid: int) -> bool:
        return rid % self.world == self.rank

    def submit(self, text: str, gender: Optional[str] = None, group: Optional[str] = None, area: Optional[str] = None,
               emotion: Optional[str] = None, sample_iteration: Optional[int] = None,
               reference_audio: Optional[str] = None, reference_text: Optional[str] = None, nfe: Optional[int] = None,
               speed: Optional[float] = None, seed: Optional[int] = None, owner: Optional[int] = None) -> Future:
        """-> Future of (int16 waveform, seconds since submit).  Requests not owned by this rank resolve to None.
        owner: the rank a front-end balancer chose for this request (`shard.dispatch_requests`); default: request
        counter modulo world."""
        cfg = self.tts.config
        with self._id_lock:
            rid = self._next_id
            self._next_id += 1
        fut: Future = Future()
        if (owner % self.world != self.rank) if owner is not None else not self.owns(rid):
            fut.set_result(None)
            return fut
        if nfe is not None and not (2 <= int(nfe) <= MAX_NFE):
            # ModelConfig caps nfe_step at 100 (model_config.py:59-62); a per-request value must not bypass that: every
            # distinct nfe costs a modulation table and a captured graph per cached batch on the device
            fut.set_exception(ValueError(f"nfe must be between 2 and {MAX_NFE}, got {nfe}"))
            return fut
        if speed is not None and not (0.1 <= speed <= 5.0):      # same range as ModelConfig.__post_init__
            fut.set_exception(ValueError(f"speed must be between 0.1 and 5.0, got {speed}"))
            return fut
        if self._closed or not self._worker.is_alive():
            fut.set_exception(RuntimeError("Speech synthesis failed: the request scheduler is closed"))
            return fut
        voice = dict(gender=gender, group=group, area=area, emotion=emotion, sample_iteration=sample_iteration,
                     reference_audio=reference_audio, reference_text=reference_text)
        self._q.put(_Request(rid, text, voice, int(nfe or cfg.nfe_step), speed,
                             int(cfg.random_seed if seed is None else seed), fut))
        return fut

    def close(self) -> None:
        self._closed = True
        self._q.put(None)
        self._worker.join(timeout=60)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # ------------------------------------------------------------------------------------------ worker side
    def _chunks_of(self, r: _Request) -> List[_Chunk]:
        v = r.voice
        ref_audio, ref_text = self.tts.model_session_manager.select_sample(
            v["gender"], v["group"], v["area"], v["emotion"], v["sample_iteration"], v["reference_audio"],
            v["reference_text"])
        inputs = self.tts._prepare_inputs(ref_audio, ref_text, r.text, speed=r.speed)
        r.n_chunks = len(inputs)
        return [_Chunk(r, i, a, ids, int(md[0])) for i, (a, ids, md, _) in enumerate(inputs)]

    def _drain(self) -> Tuple[List[_Request], bool]:
        """blocks for the first request, then collects whatever else arrives within max_wait_s"""
        first = self._q.get()
        if first is None:
            return [], True
        reqs, stop = [first], False
        deadline = time.time() + self.max_wait_s
        while True:
            left = deadline - time.time()
            try:
                nxt = self._q.get(timeout=max(left, 0.0)) if left > 0 else self._q.get_nowait()
            except queue.Empty:
                break
            if nxt is None:
                stop = True
                break
            reqs.append(nxt)
        return reqs, stop

    def _loop(self) -> None:
        cfg = self.tts.config
        eng = self.tts.model_session_manager.engine
        while True:
            reqs, stop = self._drain()
            # a request the client cancelled while it was queued is skipped; the others are RUNNING from here on
            reqs = [r for r in reqs if r.future.set_running_or_notify_cancel()]
            try:
                self._serve(reqs, cfg, eng)
            except BaseException as exc:              # nothing may take the worker down with requests in flight
                for r in reqs:
                    _resolve(r.future, error=RuntimeError(f"Speech synthesis failed: {exc}"))
            if stop:
                break
        while True:                                   # whatever was queued behind the stop sentinel
            try:
                r = self._q.get_nowait()
            except queue.Empty:
                return
            if r is not None and r.future.set_running_or_notify_cancel():
                _resolve(r.future, error=RuntimeError("Speech synthesis failed: the request scheduler is closed"))

    def _serve(self, reqs: List[_Request], cfg, eng) -> None:
        by_nfe: Dict[int, List[_Chunk]] = {}
        for r in reqs:
            try:
                for c in self._chunks_of(r):
                    by_nfe.setdefault(r.nfe, []).append(c)
            except Exception as exc:                      # same wrapping as TTSEngine.synthesize
                _resolve(r.future, error=RuntimeError(f"Speech synthesis failed: {exc}"))
        for nfe, chunks in sorted(by_nfe.items()):
            # one engine call per (micro-batch, seed): the seed is an argument of the batch call
            for seed in sorted({c.req.seed for c in chunks}):
                group = [c for c in chunks if c.req.seed == seed]
                for batch in plan_batches([c.frames for c in group], self.max_batch_chunks, self.max_batch_frames):
                    sel = [group[i] for i in batch]
                    try:
                        out = eng.synthesize_batch([c.audio for c in sel], [c.ids for c in sel],
                                                   [c.frames for c in sel], nfe=nfe, seed=seed,
                                                   chunk_keys=[c.index for c in sel])
                    except Exception as exc:
                        for c in sel:
                            _resolve(c.req.future, error=RuntimeError(f"Speech synthesis failed: {exc}"))
                        continue
                    self.batches_run += 1
                    self.chunks_run += len(sel)
                    for c, w in zip(sel, out):
                        c.req.waves[c.index] = np.array(w, copy=True).reshape(1, 1, -1)
                        self._finish(c.req, cfg)

    def _finish(self, r: _Request, cfg) -> None:
        if r.future.done() or len(r.waves) < r.n_chunks:
            return
        try:
            final = self.tts.audio_processor.concatenate_with_crossfade_improved(
                [r.waves[i] for i in range(r.n_chunks)], cfg.cross_fade_duration, cfg.sample_rate)
            _resolve(r.future, (final, time.time() - r.t0))
        except Exception as exc:
            _resolve(r.future, error=RuntimeError(f"Speech synthesis failed: {exc}"))
