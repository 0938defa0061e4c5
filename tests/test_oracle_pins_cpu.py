"""Independent cross-checks of the ORACLE's building blocks (oracle and engine shared tables and library
calls that nothing pinned to an independent source).  The oracle as a whole stays "parity unpinned" — the graphs it
restates are unreachable offline (oracle/graphs.py header) — but each piece below is pinned to a second, independent
statement of the same published definition:

  * mel filter bank  == torchaudio.functional.melscale_fbanks(513, 0, 12000, 100, 24000, None, "htk")  
  * log-mel front-end (torch.stft)     vs an explicit float64 DFT-matrix STFT with reflect padding
  * iSTFT (torch.istft)                vs an explicit float64 inverse-DFT + overlap-add + envelope division
  * attention (F.scaled_dot_product_attention)  vs an explicit float64 softmax(QK^T / sqrt(d)) V
  * interleaved RoPE (rotate-half form)         vs complex multiplication by exp(i p w_k)
  * sway-sampled time grid, sinusoidal time embedding, text position table  vs their closed forms in float64
  * ConvNeXt-V2 GRN                    vs a loop-level restatement
Everything runs on the CPU in seconds.
"""
import math

import numpy as np
import pytest
import torch

from oracle import graphs
from vietvoice_tts_b200 import artifact
from vietvoice_tts_b200.arch import FULL, TINY


@pytest.fixture(scope="module")
def tiny():
    W = artifact.make_random_weights(TINY, 9527)
    return W, graphs.OracleSessions(TINY, W)


def test_mel_filterbank_is_torchaudio_htk():
    ta = pytest.importorskip("torchaudio")
    want = ta.functional.melscale_fbanks(FULL.n_bins, float(FULL.mel_fmin), float(FULL.mel_fmax), FULL.n_mel,
                                         FULL.sample_rate, None, "htk").numpy()
    got = artifact.mel_filterbank(FULL)
    assert (FULL.n_bins, FULL.mel_fmin, FULL.mel_fmax, FULL.n_mel, FULL.sample_rate) == (513, 0.0, 12000.0, 100, 24000)
    assert got.shape == want.shape == (513, 100)
    # torchaudio evaluates the same formula in float32 (frequency points up to 12 kHz divided by band widths of
    # 50-300 Hz: ~1e-5 of rounding noise on weights in [0, 1]); ours is float64 rounded once
    assert np.abs(got - want).max() < 2e-5 and np.abs(got - want).mean() < 1e-7
    assert np.array_equal(got > 0, want > 0) or np.abs(got - want)[(got > 0) != (want > 0)].max() < 2e-5
    # and it is the table that rides in the weight blob
    assert np.array_equal(artifact.make_random_weights(TINY, 1)["pre.mel_fb"], artifact.mel_filterbank(TINY))


def _dft_mel(audio_i16, arch, fb):
    """float64, no FFT library: reflect-pad, frame, periodic Hann, DFT matrix, magnitude, filter bank, log-clamp"""
    x = audio_i16.astype(np.float64) / 32768.0
    rms = math.sqrt(float(np.mean(x * x)))
    if 0 < rms < arch.target_rms:
        x = x * (arch.target_rms / rms)
    n, hop = arch.n_fft, arch.hop
    xp = np.concatenate([x[1:n // 2 + 1][::-1], x, x[-n // 2 - 1:-1][::-1]])
    frames = 1 + (xp.size - n) // hop
    win = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n) / n)
    k = np.arange(n // 2 + 1)[:, None] * np.arange(n)[None, :]
    D = np.exp(-2j * np.pi * k / n)                                   # [bins, n]
    F = np.stack([xp[i * hop:i * hop + n] * win for i in range(frames)])
    mag = np.abs(F @ D.T)
    return np.log(np.maximum(mag @ fb.astype(np.float64), arch.mel_clamp))


def test_mel_frontend_vs_explicit_dft(tiny):
    W, S = tiny
    for n_samples, seed in ((24000, 1), (5000, 2), (777, 3)):
        audio = artifact.synthetic_prompt_pcm(n_samples, seed)
        got = S.preprocess.mel(audio).numpy()
        want = _dft_mel(audio, TINY, W["pre.mel_fb"])
        assert got.shape == want.shape == (n_samples // TINY.hop + 1, TINY.n_mel)
        assert np.abs(got - want).max() < 2e-3 and np.abs(got - want).mean() < 2e-5
    quiet = (artifact.synthetic_prompt_pcm(9000, 4).astype(np.float32) * 0.01).astype(np.int16)     # RMS branch taken
    assert np.abs(S.preprocess.mel(quiet).numpy() - _dft_mel(quiet, TINY, W["pre.mel_fb"])).max() < 5e-3


def _idft_ola(head, arch):
    nb, n, hop = arch.n_bins, arch.n_fft, arch.hop
    head = head.astype(np.float64)
    mag = np.minimum(np.exp(head[:, :nb]), arch.mag_clip)
    spec = mag * (np.cos(head[:, nb:]) + 1j * np.sin(head[:, nb:]))
    spec[:, 0] = spec[:, 0].real
    spec[:, -1] = spec[:, -1].real
    full = np.concatenate([spec, np.conj(spec[:, -2:0:-1])], axis=1)     # Hermitian extension -> [T, n]
    t = np.arange(n)
    frames = (full @ np.exp(2j * np.pi * np.outer(np.arange(n), t) / n)).real / n
    win = 0.5 - 0.5 * np.cos(2 * np.pi * t / n)
    T = head.shape[0]
    out = np.zeros(hop * (T - 1) + n)
    env = np.zeros_like(out)
    for i in range(T):
        out[i * hop:i * hop + n] += frames[i] * win
        env[i * hop:i * hop + n] += win * win
    return (out / np.where(env > 1e-11, env, 1.0))[n // 2:n // 2 + hop * (T - 1)]


def test_istft_vs_explicit_inverse_dft(tiny):
    _, S = tiny
    rng = np.random.default_rng(3)
    for T in (2, 9, 57):
        head = (rng.standard_normal((T, TINY.n_fft + 2)) * 0.7).astype(np.float32)
        head[0, 3] = 9.0                                                 # exercises the magnitude clip
        got = S.decode.wave(torch.from_numpy(head)).numpy()
        want = _idft_ola(head, TINY)
        assert got.shape == want.shape == ((T - 1) * TINY.hop,)
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())


def test_attention_vs_explicit_softmax(tiny):
    _, S = tiny
    a = TINY
    rng = np.random.default_rng(5)
    T = 75
    x_in = torch.from_numpy(rng.standard_normal((2, T, a.in_dim)).astype(np.float32))
    cos, sin = graphs.rope_tables(a, T)
    S.transformer.taps = {}
    S.transformer.velocity(x_in, 0.3, cos, sin)
    tp = S.transformer.taps
    S.transformer.taps = None
    q, k, v = (tp[n].double().numpy() for n in ("q0", "k0", "v0"))
    want = np.empty_like(q)
    for b in range(2):
        for h in range(a.heads):
            sl = slice(h * a.head_dim, (h + 1) * a.head_dim)
            s = q[b][:, sl] @ k[b][:, sl].T / math.sqrt(a.head_dim)
            p = np.exp(s - s.max(axis=1, keepdims=True))
            want[b][:, sl] = (p / p.sum(axis=1, keepdims=True)) @ v[b][:, sl]
    assert np.abs(tp["attn0"].numpy() - want).max() < 5e-5


def test_rope_is_complex_rotation_of_interleaved_pairs():
    a = TINY
    T = 40
    cos, sin = graphs.rope_tables(a, T)
    rng = np.random.default_rng(6)
    x = rng.standard_normal((T, a.head_dim))
    got = (torch.from_numpy(x).float() * cos + graphs._rotate_half_interleaved(torch.from_numpy(x).float()) * sin).numpy()
    w = a.rope_theta ** (-np.arange(0, a.head_dim, 2) / a.head_dim)
    z = (x[:, 0::2] + 1j * x[:, 1::2]) * np.exp(1j * np.arange(T)[:, None] * w[None, :])
    want = np.stack([z.real, z.imag], axis=-1).reshape(T, a.head_dim)
    assert np.abs(got - want).max() < 1e-5


def test_time_grid_time_embedding_and_position_table(tiny):
    W, S = tiny
    a = TINY
    for nfe in (2, 16, 32, 64):
        t = graphs.time_grid(a, nfe).numpy()
        u = np.arange(nfe) / (nfe - 1)
        assert np.abs(t - (u + a.sway * (np.cos(np.pi / 2 * u) - 1 + u))).max() < 1e-12
        assert t[0] == 0.0 and abs(t[-1] - 1.0) < 1e-12 and np.all(np.diff(t) > 0)
    half = a.time_freq_dim // 2
    tt = 0.37
    f = np.exp(-math.log(10000.0) * np.arange(half) / (half - 1))
    e = np.concatenate([np.sin(1000 * tt * f), np.cos(1000 * tt * f)])
    w1, b1 = W["dit.time.l1.w"].astype(np.float64), W["dit.time.l1.b"].astype(np.float64)
    w2, b2 = W["dit.time.l2.w"].astype(np.float64), W["dit.time.l2.b"].astype(np.float64)
    h = w1 @ e + b1
    want = w2 @ (h / (1 + np.exp(-h))) + b2
    assert np.abs(S.transformer.time_embed(tt).numpy() - want).max() < 1e-4
    td = a.text_dim
    p = np.arange(7)[:, None] * (10000.0 ** (-np.arange(0, td, 2)[: td // 2] / td))[None, :]
    assert np.abs(S.preprocess.pos_table[:7].numpy() - np.concatenate([np.cos(p), np.sin(p)], axis=1)).max() < 1e-6


def test_text_convnext_grn_loop_level(tiny):
    """One ConvNeXt-V2 block of the text embedding restated with explicit loops (depthwise conv taps, LayerNorm,
    exact GELU, GRN over the TIME axis) in float64."""
    W, S = tiny
    a = TINY
    rng = np.random.default_rng(8)
    T = 23
    ids = torch.from_numpy(rng.integers(1, a.vocab + 1, size=T))
    got = S.preprocess.text_embed(ids, T).double().numpy()
    x = W["pre.text_embed"].astype(np.float64)[ids.numpy()] + S.preprocess.pos_table[:T].double().numpy()
    erf = np.vectorize(math.erf)
    for i in range(a.text_layers):
        g = lambda n: W[f"pre.text_blocks.{i}.{n}"].astype(np.float64)
        h = np.zeros_like(x)
        for t in range(T):
            for k in range(7):
                if 0 <= t + k - 3 < T:
                    h[t] += g("dw.w")[:, k] * x[t + k - 3]
        h += g("dw.b")
        mu, var = h.mean(axis=1, keepdims=True), h.var(axis=1, keepdims=True)
        h = (h - mu) / np.sqrt(var + a.ln_eps) * g("ln.g") + g("ln.b")
        h = h @ g("pw1.w").T + g("pw1.b")
        h = 0.5 * h * (1 + erf(h / math.sqrt(2)))
        gx = np.sqrt((h * h).sum(axis=0, keepdims=True))
        h = g("grn.g") * (h * (gx / (gx.mean() + 1e-6))) + g("grn.b") + h
        x = x + h @ g("pw2.w").T + g("pw2.b")
    assert np.abs(got - x).max() < 2e-4 * max(1.0, np.abs(x).max())
