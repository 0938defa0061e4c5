"""Request scheduler: the packing function and the request -> chunk -> micro-batch -> waveform
bookkeeping, with a stand-in engine (no GPU).  The GPU counterpart is tests/test_scheduler_gpu.py."""
import threading

import numpy as np
import pytest

from vietvoice_tts_b200.host.scheduler import RequestScheduler, plan_batches


def test_plan_batches_is_a_partition_within_limits():
    rng = np.random.default_rng(0)
    frames = [int(x) for x in rng.integers(600, 1800, size=37)]
    batches = plan_batches(frames, max_chunks=8, max_frames=8 * 1500)
    flat = sorted(i for b in batches for i in b)
    assert flat == list(range(37))                                   # every chunk exactly once
    for b in batches:
        assert 1 <= len(b) <= 8
        assert sum(frames[i] for i in b) <= 8 * 1500 or len(b) == 1
    lens = [[frames[i] for i in b] for b in batches]
    assert all(min(lens[k]) >= max(lens[k + 1]) for k in range(len(lens) - 1))   # longest first, similar together
    assert plan_batches([], 8, 1000) == []
    assert plan_batches([5000], 8, 1000) == [[0]]                     # oversized chunk still gets a batch


class _FakeAudio:
    def concatenate_with_crossfade_improved(self, waves, fade, sr):
        return np.concatenate([w.reshape(-1) for w in waves])


class _FakeManager:
    def __init__(self, engine):
        self.engine = engine

    def select_sample(self, *a):
        return b"prompt", "ref"


class _FakeEngine:
    def __init__(self):
        self.calls = []

    def synthesize_batch(self, audios, ids, frames, nfe=0, seed=0, chunk_keys=None):
        self.calls.append((list(frames), nfe, seed, list(chunk_keys)))
        # waveform encodes (first id, chunk key, nfe, seed) so the test can check the routing
        return [np.full(4, int(i[0, 0]) * 1000 + k * 100 + nfe + seed, dtype=np.int32) for i, k in zip(ids, chunk_keys)]


class _FakeConfig:
    nfe_step = 32
    random_seed = 7
    cross_fade_duration = 0.1
    sample_rate = 24000


class _FakeTTS:
    def __init__(self):
        self.config = _FakeConfig()
        self.audio_processor = _FakeAudio()
        self.model_session_manager = _FakeManager(_FakeEngine())

    def _prepare_inputs(self, ref_audio, ref_text, text, speed=None):
        n = int(text.split(":")[1])                                  # "id:nchunks"
        rid = int(text.split(":")[0])
        return [(np.zeros((1, 1, 8), np.int16), np.full((1, 3), rid, np.int32),
                 np.array([700 + 100 * c + (0 if speed is None else int(speed * 10))], np.int64), np.array([0], np.int32))
                for c in range(n)]


def test_requests_are_batched_by_nfe_and_routed_back():
    tts = _FakeTTS()
    with RequestScheduler(tts, max_batch_chunks=4, max_batch_frames=10 ** 6, max_wait_s=0.2) as sch:
        futs = [sch.submit(f"{i}:{1 + i % 3}", nfe=(16 if i % 2 else None), speed=(1.5 if i == 4 else None))
                for i in range(6)]
        res = [f.result(timeout=30) for f in futs]
    eng = tts.model_session_manager.engine
    assert sum(len(c[0]) for c in eng.calls) == sum(1 + i % 3 for i in range(6)) == sch.chunks_run
    assert all(len(c[0]) <= 4 for c in eng.calls)
    assert {c[1] for c in eng.calls} == {16, 32}                     # one graph per nfe: never mixed in a batch
    for i, (wave, secs) in enumerate(res):
        nfe = 16 if i % 2 else 32
        n = 1 + i % 3
        want = np.concatenate([np.full(4, i * 1000 + k * 100 + nfe + 7, np.int32) for k in range(n)])
        assert np.array_equal(wave, want) and secs >= 0                # chunks back in order, keyed 0..n-1, own nfe
    assert any(4 == len(c[0]) for c in eng.calls)                    # chunks of different requests share a batch
    assert any(715 in c[0] for c in eng.calls)                       # per-request speed reached _prepare_inputs


def test_rank_ownership_and_validation():
    tts = _FakeTTS()
    with RequestScheduler(tts, rank=1, world=2, max_wait_s=0.01) as sch:
        a = sch.submit("0:1")          # request 0 -> rank 0
        b = sch.submit("1:1")          # request 1 -> rank 1 (this one)
        assert a.result(timeout=10) is None
        assert b.result(timeout=10)[0].shape == (4,)
        assert sch.submit("2:1").result(timeout=10) is None
        bad = sch.submit("3:1", speed=9.0)       # request 3 -> this rank; speed outside ModelConfig's range
        with pytest.raises(ValueError):
            bad.result(timeout=10)
        # an explicit owner (front-end balancer) overrides the counter rule
        assert sch.submit("4:1", owner=0).result(timeout=10) is None
        assert sch.submit("5:1", owner=1).result(timeout=10)[0].shape == (4,)
        assert sch.submit("6:1", owner=3).result(timeout=10)[0].shape == (4,)     # owner taken modulo world


def test_dispatch_requests_balances_and_is_deterministic():
    from vietvoice_tts_b200.shard import dispatch_requests
    rng = np.random.default_rng(5)
    costs = [float(rng.integers(20, 300) * rng.choice([15, 31, 63])) for _ in range(384)]
    for world in (1, 2, 8):
        own = dispatch_requests(costs, world)
        assert own == dispatch_requests(costs, world) and set(own) <= set(range(world))
        load = [sum(c for c, o in zip(costs, own) if o == r) for r in range(world)]
        rr = [sum(costs[i] for i in range(len(costs)) if i % world == r) for r in range(world)]
        assert max(load) <= max(rr) + 1e-9                       # never worse than round robin on this stream
        assert max(load) / (sum(load) / world) < 1.03            # online least-loaded: within a few % of the mean
    assert dispatch_requests([], 4) == [] and dispatch_requests([1.0, 1.0, 1.0], 2) == [0, 1, 0]


def test_submit_is_thread_safe():
    tts = _FakeTTS()
    out = {}
    with RequestScheduler(tts, max_wait_s=0.02) as sch:
        def work(k):
            out[k] = sch.submit(f"{k}:2").result(timeout=30)
        ts = [threading.Thread(target=work, args=(k,)) for k in range(16)]
        [t.start() for t in ts]
        [t.join() for t in ts]
    assert len(out) == 16 and all(v[0].shape == (8,) for v in out.values())
    assert tts.model_session_manager.engine.calls and sch.chunks_run == 32


def test_worker_survives_cancels_engine_errors_and_close():
    """a cancelled future made set_result raise InvalidStateError inside the worker, which killed the
    daemon thread and left every later request hanging; requests queued behind close()'s sentinel never resolved;
    a client-supplied nfe had no upper bound."""
    import time

    class _SlowEngine(_FakeEngine):
        def synthesize_batch(self, *a, **k):
            if k.get("nfe") == 17:
                raise RuntimeError("device fell over")
            time.sleep(0.05)
            return super().synthesize_batch(*a, **k)

    tts = _FakeTTS()
    tts.model_session_manager.engine = _SlowEngine()
    sch = RequestScheduler(tts, max_wait_s=0.01)
    first = sch.submit("0:1")
    queued = [sch.submit(f"{i}:1") for i in range(1, 6)]
    cancelled = [f for f in queued if f.cancel()]               # still queued behind the slow first batch
    assert first.result(timeout=10)[0].shape == (4,)
    for f in queued:
        if f not in cancelled:
            assert f.result(timeout=10)[0].shape == (4,)
    boom = sch.submit("6:2", nfe=17)                            # engine error: wrapped, the worker carries on
    with pytest.raises(RuntimeError, match="Speech synthesis failed: device fell over"):
        boom.result(timeout=10)
    assert sch.submit("7:1").result(timeout=10)[0].shape == (4,)   # the worker is still alive
    assert sch._worker.is_alive()
    for bad in (1, 101, 100000):
        with pytest.raises(ValueError, match="nfe must be between 2 and 100"):
            sch.submit("8:1", nfe=bad).result(timeout=10)
    sch.close()
    assert not sch._worker.is_alive()
    with pytest.raises(RuntimeError, match="scheduler is closed"):
        sch.submit("9:1").result(timeout=10)                     # resolved at once instead of queueing forever
