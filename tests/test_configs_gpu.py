"""BASELINE.json configs 3-5 as parity cases on the B200 (configs[1] is the bench line, configs[0] the CPU leg).

  cfg 3  long text chunked by sentence, ragged T in one batch      -> each chunk vs its own oracle run (TINY), and at
                                                                    FULL size: every chunk alone == inside the batch
  cfg 4  voice clone, long conditioning context (T_ref 851, T 1780) -> shapes, determinism and batch invariance at FULL
                                                                    size (the oracle needs minutes there), oracle
                                                                    parity at the same prompt/target ratio on TINY
  cfg 5  request stream with NFE 16 / 32 / 64                       -> all nfe-1 Euler steps vs the oracle built for
                                                                    that nfe (TINY): rel-L2 <= 2e-2, SNR >= 25 dB
Tolerances are the ones of tests/test_engine_gpu.py."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL, TINY
from vietvoice_tts_b200.engine import Engine
from oracle.graphs import OracleSessions


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / (np.linalg.norm(b) + 1e-30))


def snr_db(x, ref):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return 10 * np.log10(np.sum(ref ** 2) / (np.sum((x - ref) ** 2) + 1e-30))


@pytest.fixture(scope="module")
def tiny():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W = artifact.make_random_weights(TINY, 9527)
    eng = Engine.from_weights(TINY, W, device=0)
    yield eng, W
    eng.close()


@pytest.fixture(scope="module")
def full():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    W = artifact.make_random_weights(FULL, 9527)
    eng = Engine.from_weights(FULL, W, device=0)
    del W
    yield eng
    eng.close()


@pytest.mark.parametrize("nfe", [16, 32, 64])
def test_cfg5_nfe_sweep_matches_oracle(tiny, nfe):
    """per-request NFE (15 / 31 / 63 Euler steps, each its own time grid + CUDA graph) against the oracle"""
    eng, W = tiny
    ora = OracleSessions(TINY, W, nfe=nfe)
    n_samples, T = 20000, 20000 // 256 + 1 + 75
    rng = np.random.default_rng(nfe)
    audio = artifact.synthetic_prompt_pcm(n_samples, nfe)
    ids = rng.integers(0, TINY.vocab, size=(1, 33)).astype(np.int32)
    noise = rng.standard_normal((1, T, TINY.n_mel)).astype(np.float32)
    wave, x, _, pre = ora.synthesize_chunk(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    b = eng.batch([T])
    b.preprocess(0, audio, ids, noise)
    b.sample(nfe=nfe)
    got = b.get(0, "noise")
    assert rel(got, x[0]) < 2e-2, rel(got, x[0])
    pcm = b.decode(0)
    assert pcm.shape[0] == wave.size == (T - int(pre[7][0]) - 1) * 256
    assert snr_db(pcm, wave.reshape(-1)) > 25.0
    b.close()


def test_cfg4_long_prompt_ratio_matches_oracle(tiny):
    """voice clone: the prompt is ~48 % of the frames (9.07 s prompt / 9.9 s target) and the text is prompt text +
    target text without separator (core/tts_engine.py:121-122)"""
    eng, W = tiny
    ora = OracleSessions(TINY, W)
    n_samples = 54400                       # T_ref 213
    T = n_samples // 256 + 1 + 232
    rng = np.random.default_rng(44)
    audio = artifact.synthetic_prompt_pcm(n_samples, 44)
    ids = rng.integers(0, TINY.vocab, size=(1, 150)).astype(np.int32)
    noise = rng.standard_normal((1, T, TINY.n_mel)).astype(np.float32)
    wave, x, _, pre = ora.synthesize_chunk(audio.reshape(1, 1, -1), ids, np.array([T], dtype=np.int64), noise)
    b = eng.batch([T])
    assert b.preprocess(0, audio, ids, noise) == 213
    b.sample()
    assert rel(b.get(0, "noise"), x[0]) < 2e-2
    assert snr_db(b.decode(0), wave.reshape(-1)) > 25.0
    b.close()


def test_cfg4_full_size_voice_clone_properties(full):
    """FULL architecture at the cfg-4 size: prompt 9.0704 s -> N 217 689 -> T_ref 851, target 9.9 s -> T 1780"""
    eng = full
    n_samples, T = 217689, 851 + 929
    rng = np.random.default_rng(4)
    audios = [artifact.synthetic_prompt_pcm(n_samples, 70 + i) for i in range(2)]
    ids = [rng.integers(0, FULL.vocab, size=320).astype(np.int32) for _ in range(2)]
    a = eng.synthesize_batch(audios, ids, [T, T], nfe=6, seed=5, chunk_keys=[0, 1])
    b = eng.synthesize_batch(audios, ids, [T, T], nfe=6, seed=5, chunk_keys=[0, 1])
    solo = eng.synthesize_batch(audios[1:], ids[1:], [T], nfe=6, seed=5, chunk_keys=[1])
    for x, y in zip(a, b):
        assert x.dtype == np.int16 and x.shape[0] == (T - 851 - 1) * 256
        assert np.array_equal(x, y)                      # replay is bit-identical
        assert np.abs(x.astype(np.int32)).max() > 0
    assert snr_db(solo[0], a[1]) > 25.0                  # independent of what else is in the batch


def test_cfg3_ragged_chunks_full_size_batch_invariance(full):
    """long text -> sentence chunks with targets of 3..13 s behind one 6 s prompt: T in [845, 1782] in ONE batch;
    every chunk must come out as if it had been synthesised alone, and chunk order must not matter (noise is keyed
    on (seed, chunk index))"""
    eng = full
    rng = np.random.default_rng(3)
    Ts = [845, 1782, 1203, 1501, 977]
    audio = artifact.synthetic_prompt_pcm(144000, 9)
    ids = [rng.integers(0, FULL.vocab, size=int(100 + 0.18 * (t - 563))).astype(np.int32) for t in Ts]
    keys = list(range(len(Ts)))
    whole = eng.synthesize_batch([audio] * len(Ts), ids, Ts, nfe=5, seed=7, chunk_keys=keys)
    for t, w in zip(Ts, whole):
        assert w.shape[0] == (t - 563 - 1) * 256
    perm = [3, 0, 4, 2, 1]
    shuffled = eng.synthesize_batch([audio] * len(Ts), [ids[i] for i in perm], [Ts[i] for i in perm], nfe=5, seed=7,
                                    chunk_keys=[keys[i] for i in perm])
    for j, i in enumerate(perm):
        assert snr_db(shuffled[j], whole[i]) > 25.0
    for i in (0, 1):
        solo = eng.synthesize_batch([audio], [ids[i]], [Ts[i]], nfe=5, seed=7, chunk_keys=[i])
        assert snr_db(solo[0], whole[i]) > 25.0
