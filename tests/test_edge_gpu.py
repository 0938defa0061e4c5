"""Edge cases of the C-ABI path on the B200: empty text, the shortest legal chunk, bad arguments (status code + message,
never a crash — the error contract of /root/reference/vietvoicetts/core/tts_engine.py:256-257), and the bounded
batch cache behind `vv_synthesize_batch` (a request stream produces a new frame-count key per micro-batch)."""
import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY
from vietvoice_tts_b200.engine import Engine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def setup():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from oracle.graphs import OracleSessions
    torch.set_num_threads(4)
    W = artifact.make_random_weights(TINY, 9527)
    eng = Engine.from_weights(TINY, W)
    yield eng, OracleSessions(TINY, W)
    eng.close()


def _snr(a, b):
    a, b = a.astype(np.float64), b.astype(np.float64)
    return 10 * np.log10((b ** 2).sum() / max(((a - b) ** 2).sum(), 1e-30))


def test_empty_text_is_all_filler(setup):
    """text_to_indices("") is a [1, 0] array (core/text_processor.py:30-37): every text row is the filler token"""
    eng, S = setup
    audio = artifact.synthetic_prompt_pcm(8000, 2)
    T = 8000 // 256 + 1 + 40
    rng = np.random.default_rng(1)
    noise = rng.standard_normal((T, TINY.n_mel)).astype(np.float32)
    ids = np.zeros((0,), dtype=np.int32)
    got = eng.synthesize_batch([audio], [ids], [T], noises=[noise])[0]
    ref = S.synthesize_chunk(audio.reshape(1, 1, -1), ids.reshape(1, 0), np.array([T], dtype=np.int64),
                             noise=noise[None])
    ref = np.asarray(ref[0]).reshape(-1)
    assert got.shape == ref.shape == ((T - (8000 // 256 + 1) - 1) * 256,)
    assert _snr(got, ref) > 25.0          # dB, same bar as tests/test_engine_gpu.py


def test_shortest_chunk_and_zero_length_output(setup):
    """T = ref_len + 1 leaves (T - ref_len - 1) * hop = 0 samples: an empty waveform, not an error"""
    eng, _ = setup
    audio = artifact.synthetic_prompt_pcm(6000, 3)
    ref_len = 6000 // 256 + 1
    ids = np.arange(5, dtype=np.int32)
    out = eng.synthesize_batch([audio, audio], [ids, ids], [ref_len + 1, ref_len + 2])
    assert out[0].size == 0 and out[1].size == 256


def test_bad_arguments_return_errors_not_crashes(setup):
    eng, _ = setup
    audio = artifact.synthetic_prompt_pcm(6000, 3)
    ids = np.arange(5, dtype=np.int32)
    T = 6000 // 256 + 1 + 20
    with pytest.raises(RuntimeError, match="out of vocabulary"):
        eng.synthesize_batch([audio], [np.array([TINY.vocab], dtype=np.int32)], [T])
    with pytest.raises(RuntimeError, match="out of vocabulary"):
        eng.synthesize_batch([audio], [np.array([-1], dtype=np.int32)], [T])
    with pytest.raises(RuntimeError, match="too short"):
        eng.synthesize_batch([audio[:100]], [ids], [T])
    with pytest.raises(RuntimeError, match="out of range"):
        eng.synthesize_batch([audio], [ids], [1])
    with pytest.raises(RuntimeError, match="out of range"):
        eng.synthesize_batch([audio], [ids], [10 ** 6])
    with pytest.raises(RuntimeError, match="pcm_capacity"):
        eng.synthesize_batch([audio], [ids], [T], pcm_out=[np.empty(8, dtype=np.int16)])
    # the engine is still usable after every failure
    assert eng.synthesize_batch([audio], [ids], [T])[0].size == (T - (6000 // 256 + 1) - 1) * 256


def test_concurrent_callers_share_one_engine(setup):
    """the reference's REST layer runs requests on worker threads against one engine (api/tts_engine.py:79-87):
    four threads call the C ABI on the same handle without any Python-side lock; results equal the serial ones"""
    import threading
    eng, _ = setup
    audio = artifact.synthetic_prompt_pcm(6000, 3)
    ref = 6000 // 256 + 1
    jobs = [(np.arange(3 + j, dtype=np.int32), [ref + 30 + 7 * j, ref + 50 + 5 * j], 100 + j) for j in range(8)]

    def run(j):
        ids, T, key = jobs[j]
        return eng.synthesize_batch([audio, audio], [ids, ids], T, chunk_keys=[key, key + 1000])

    serial = [run(j) for j in range(len(jobs))]
    got = [None] * len(jobs)
    errs = []

    def worker(t):
        try:
            for j in range(t, len(jobs), 4):
                got[j] = run(j)
        except Exception as ex:            # pragma: no cover
            errs.append(ex)

    th = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs
    for a, b in zip(serial, got):
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


_CACHE_SCRIPT = r"""
import sys, numpy as np, torch
sys.path.insert(0, %r)
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY
from vietvoice_tts_b200.engine import Engine
eng = Engine.from_weights(TINY, artifact.make_random_weights(TINY, 9527))
audio = artifact.synthetic_prompt_pcm(6000, 3)
ids = np.arange(9, dtype=np.int32)
ref = 6000 // 256 + 1
first = eng.synthesize_batch([audio] * 4, [ids] * 4, [ref + 300] * 4)
free = []
for k in range(24):                      # 24 distinct keys through a cache of 2
    T = ref + 40 + 16 * (23 - k)           # shrinking batches: any growth in use is a leak, not the batch itself
    eng.synthesize_batch([audio] * 4, [ids] * 4, [T, T + 1, T + 2, T + 3])
    free.append(torch.cuda.mem_get_info()[0])
again = eng.synthesize_batch([audio] * 4, [ids] * 4, [ref + 300] * 4)   # evicted long ago: rebuilt, same bits
assert all(np.array_equal(a, b) for a, b in zip(first, again))
print("FREE", free[3], min(free[4:]))
"""


def test_batch_cache_is_bounded():
    """24 different batch shapes with VVB200_BATCH_CACHE=2: device memory in use stops growing after the cache is
    full, and a shape that was evicted gives bit-identical output when it comes back"""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    env = dict(os.environ, VVB200_BATCH_CACHE="2")
    r = subprocess.run([sys.executable, "-c", _CACHE_SCRIPT % ROOT], env=env, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("FREE")][-1].split()
    after_fill, lowest_later = int(line[1]), int(line[2])
    # batches shrink over the run, so with two cached batches the free memory must not drop after the cache filled
    # up; an unbounded cache would be down by 20 batches' worth (hundreds of MB even at TINY size)
    assert after_fill - lowest_later < 32 << 20
