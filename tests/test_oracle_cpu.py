"""The oracle itself, on CPU: I/O contract of the three graphs (arity, order, dtypes, shapes — the ABI the reference
binds by position, /root/reference/vietvoicetts/core/tts_engine.py:133-187, :229-230), the loop count, and a
regression pin against committed vectors (tests/golden/oracle_tiny.npz)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")

from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import TINY, FULL
from oracle.graphs import OracleSessions, time_grid

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def sessions():
    torch.set_num_threads(2)
    return OracleSessions(TINY, artifact.make_random_weights(TINY, 9527))


def test_session_io_contract(sessions):
    S = sessions
    assert len(S.preprocess.input_names) == 3 and len(S.preprocess.output_names) == 8
    assert len(S.transformer.input_names) == 8 and len(S.transformer.output_names) == 2
    assert len(S.decode.input_names) == 2 and len(S.decode.output_names) == 1
    n, T = 10000, 10000 // 256 + 1 + 30
    audio = artifact.synthetic_prompt_pcm(n, 3).reshape(1, 1, -1)
    ids = np.arange(20, dtype=np.int32)[None] % TINY.vocab
    pre = S.preprocess.run(audio, ids, np.array([T], dtype=np.int64))
    noise, cq, sq, ck, sk, cat_c, cat_u, ref_len = pre
    assert noise.shape == (1, T, 100) and noise.dtype == np.float32
    assert cq.shape == sq.shape == (1, T, 64) and ck.shape == sk.shape == (1, 64, T)
    assert cat_c.shape == cat_u.shape == (1, T, TINY.cond_dim)
    assert ref_len.dtype == np.int64 and int(ref_len[0]) == n // 256 + 1        # == host ref_audio_len (tts_engine.py:55)
    assert np.all(cat_u[0, :, :100] == 0) and np.all(cat_c[0, int(ref_len[0]):, :100] == 0)
    x, ts = S.transformer.run(noise, cq, sq, ck, sk, cat_c, cat_u, np.array([0], dtype=np.int32))
    assert x.shape == noise.shape and ts.dtype == np.int32 and int(ts[0]) == 1
    wave = S.decode.run(x, ref_len)[0]
    assert wave.dtype == np.int16 and wave.reshape(-1).shape[0] == (T - int(ref_len[0]) - 1) * 256


def test_time_grid():
    t = time_grid(FULL).numpy()
    assert t.shape == (32,) and t[0] == 0.0 and abs(t[-1] - 1.0) < 1e-12 and np.all(np.diff(t) > 0)
    assert abs(t[1] - (1 / 31 - (np.cos(np.pi / 2 / 31) - 1 + 1 / 31))) < 1e-12      # sway s = -1


def test_noise_is_injected_or_seeded(sessions):
    n, T = 9000, 9000 // 256 + 1 + 10
    audio = artifact.synthetic_prompt_pcm(n, 3).reshape(1, 1, -1)
    ids = np.zeros((1, 4), dtype=np.int32)
    z = np.random.default_rng(0).standard_normal((1, T, 100)).astype(np.float32)
    assert np.array_equal(sessions.preprocess.run(audio, ids, np.array([T]), z)[0], z)
    a = sessions.preprocess.run(audio, ids, np.array([T]))[0]
    b = sessions.preprocess.run(audio, ids, np.array([T]))[0]
    assert not np.array_equal(a, b)          # the session RNG advances per call, like ORT's


def test_regression_pin(sessions):
    g = np.load(os.path.join(G, "oracle_tiny.npz"))
    T = int(g["T"][0])
    wave, x, steps, pre = sessions.synthesize_chunk(g["audio"].reshape(1, 1, -1), g["ids"], np.array([T]), g["noise"], True)
    assert len(steps) == TINY.nfe - 1                                  # nfe_step - 1 calls (tts_engine.py:157)
    assert int(pre[7][0]) == int(g["ref_signal_len"][0])
    assert np.abs(pre[5] - g["cat_mel_text"].astype(np.float32)).max() < 2e-2
    assert np.abs(steps[0] - g["step1"]).max() < 1e-3
    assert np.abs(x - g["final"]).max() < 5e-3
    ref = g["wave"].reshape(-1).astype(np.float64)
    d = wave.reshape(-1).astype(np.float64) - ref
    assert 10 * np.log10((ref ** 2).sum() / ((d ** 2).sum() + 1e-30)) > 40.0
